cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q -s > gpurun_out/r1w_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r1w_tests.log;  grep -a "impulse form\|lifted" gpurun_out/r1w_tests.log | cut -c1-300
python bench.py --steps 110 --warmup 5 --no-cpu-baseline > gpurun_out/r1w_bench.json 2> gpurun_out/r1w_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r1w_bench.json')); print(round(d['value']), d['step_ms_quantiles'], d['e2e']['value'], d['episode_stats']); print(d['step_ms_first_60'])
"
PASSES=1 LAST=2 python tools/timeline.py 51 2>&1 | grep -v "branch L" | tail -20
PASSES=1 LAST=1 python tools/timeline.py 30 2>&1 | grep -v "branch L" | tail -10
