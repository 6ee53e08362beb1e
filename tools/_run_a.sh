cd $GRAFT_REPO_ROOT
python bench.py --steps 2 --warmup 47 --no-graph --no-cpu-baseline --e2e-steps 1 > gpurun_out/r1z_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_heavy_solve2|k_heavy_rows' -s 20960 -c 4 -o gpurun_out/r1z_wave python bench.py --steps 2 --warmup 47 --no-graph --no-cpu-baseline --e2e-steps 1 > gpurun_out/r1z_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r1z_ncu.log | cut -c1-200
