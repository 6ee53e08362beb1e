cd $GRAFT_REPO_ROOT
python bench.py --steps 110 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2a_bench.json')); print(round(d['value']), d['step_ms_quantiles'], d['e2e']['value'], d['episode_stats']); print(d['step_ms_first_60'])
"
XARM_TRACE_STAGES=1 python bench.py --steps 40 --warmup 3 --no-graph --no-cpu-baseline --e2e-steps 1 2>&1 | grep "xarm lists" | tail -12
PASSES=1 LAST=1 python tools/timeline.py 30 2>&1 | grep -v "branch L" | tail -9
