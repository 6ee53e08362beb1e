#!/bin/bash
# Development: short staggered bench lines of the other four BASELINE configs (one summary line each; the JSON lines go to $OUT).
# usage: [STEPS=60] [OUT=gpurun_out/x.jsonl] tools/bench_tasks.sh [task ...]      env knobs (XARM_*) pass through
tasks=${@:-reach stack_tower push_with_door handover}
for t in $tasks; do
  line=$(python bench.py --task $t --steps ${STEPS:-60} --warmup 5 --sync-steps 0 --no-cpu-baseline --e2e-steps 3 2>/dev/null | tail -1)
  [ -n "$OUT" ] && echo "$line" >> $OUT
  echo "$line" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$t', '| %.3f M env-steps/s | %.2f ms/step | branch ms' % (d['value']/1e6, d['ms_per_step']), {k: round(v,2) for k,v in r['kernel_ms_per_step_by_branch'].items()}, '| heavy frac', {k: (round(v,4) if v is not None else None) for k,v in r['heavy_env_substep_fraction'].items()}, '| kernels', [(k['kernel'], round(k['avg_us']), round(k['share_of_kernel_time'],2)) for k in r['kernels'][:4]], '| eps', d['episode_stats'])
"
done
