set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python scripts/exp_growth.py 1024 8192 500 12 net 2>&1 | tail -14
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -c 3000 gpurun_out/bench_r1.json; tail -5 gpurun_out/bench_r1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --sims 120 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_vn_conv|k_vn_fc|k_select_expand|k_backup" -s 200 -c 4 -o gpurun_out/prof_r1 python bench.py --steps 1 --warmup 1 --sims 60 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
