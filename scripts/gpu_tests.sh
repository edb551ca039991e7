set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 700 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --sims 150 --no-cpu-baseline --no-secondary > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tc_conv|k_tc_fc|k_select_expand|k_backup" -s 300 -c 4 -o gpurun_out/prof_r1 python bench.py --steps 1 --warmup 1 --sims 100 --no-cpu-baseline --no-secondary > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -4
