#!/bin/bash
# Development helper: build an A/B variant of the library (one task only) into gym_xarm_b200/_lib/variants/.
# usage: tools/build_variant.sh <name> [extra nvcc -D flags...]      (XARM_TASK=<id> selects the task, default 1)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
name=$1; shift
mkdir -p "$ROOT/gym_xarm_b200/_lib/variants"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -shared \
  -DXARM_ONLY_TASK=${XARM_TASK:-1} "$@" -Xptxas -v -o "$ROOT/gym_xarm_b200/_lib/variants/lib_$name.so" "$ROOT/gym_xarm_b200/csrc/xarm_lib.cu" 2>&1 \
  | grep -E "error|k_stepI" -A1 | grep -E "error|Used" | head -3
