#!/bin/bash
# Development: short staggered bench runs under different XARM_* knobs (one JSON summary line each).
# usage: tools/sweep_knobs.sh "XARM_RESERVE_SMS=48" "XARM_RESERVE_SMS=48 XARM_SETUP_BPS=2" ...
for kv in "$@"; do
  out=$(env $kv python bench.py --steps ${STEPS:-60} --warmup 5 --sync-steps 0 --no-cpu-baseline --e2e-steps 3 2>/dev/null | tail -1)
  echo "$out" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$kv', '| value %.3f M | ms %.2f | branch ms' % (d['value']/1e6, d['ms_per_step']), {k: round(v,2) for k,v in r['kernel_ms_per_step_by_branch'].items()}, '| p50 %.2f' % d['step_ms_quantiles']['p50'])
"
done
