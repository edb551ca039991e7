# round-end measurement on the GPU box: bench line, ncu launch list of the bench command, one ncu --set full capture of the top kernels
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 420 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; python scripts/show_bench.py gpurun_out/bench_r1_final.json; tail -2 gpurun_out/bench_r1_final.err
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 600 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --sims 150 --no-cpu-baseline --no-secondary > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/launches_r1.csv | cut -c1-200
timeout 230 ncu --set full --clock-control none --import-source on -k regex:"k_tc_conv|k_tc_fc|k_select_expand|k_backup" -s 300 -c 4 -o gpurun_out/prof_r1 python bench.py --steps 1 --warmup 1 --sims 100 --no-cpu-baseline --no-secondary > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -6
