#!/bin/bash
# Round-end measurement set (one GPU): bench lines of all five BASELINE configs, the CPU arm, the ncu traffic pass of one step
# (tools/ncu_traffic.py) and an `ncu --set full` capture of the main branch's substep kernels at full size, summarised on the box
# (the reports exceed gpurun's return limit).  Outputs: gpurun_out/${TAG}_*.   usage: TAG=r2y tools/final_round.sh
TAG=${TAG:-r2y}; O=gpurun_out
python bench.py > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench_default.log
for t in reach stack_tower push_with_door; do python bench.py --task $t --steps 200 --sync-steps 0 --e2e-steps 40 > $O/${TAG}_bench_$t.json 2>/dev/null; done
python bench.py --task handover --steps 100 --sync-steps 0 --e2e-steps 20 > $O/${TAG}_bench_handover.json 2>/dev/null
python bench.py --impl reference --steps 20 --warmup 5 > $O/${TAG}_bench_reference_arm.json 2>/dev/null
ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
    --log-file /tmp/traffic_raw.csv python tools/ncu_traffic.py run > $O/${TAG}_traffic_run.txt 2>&1
python tools/ncu_traffic.py parse /tmp/traffic_raw.csv > $O/${TAG}_traffic_summary.txt 2>&1
cp profiles/traffic_pick_and_place.json $O/traffic_pick_and_place.json
grep -v "^==" /tmp/traffic_raw.csv | gzip > $O/${TAG}_launches_one_step.csv.gz
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"k_heavy_fused|k_pipe_light_lat|k_pipe_setup" \
    --launch-skip 210 --launch-count 6 -o /tmp/main python tools/ncu_traffic.py run > $O/${TAG}_ncu_full_run.txt 2>&1
python tools/ncu_summary.py /tmp/main.ncu-rep > $O/${TAG}_ncu_full_main_kernels_summary.txt 2>&1
ncu -i /tmp/main.ncu-rep --page raw --csv 2>/dev/null | gzip > $O/${TAG}_ncu_full_main_kernels_raw.csv.gz
ls -la $O | tail -20
