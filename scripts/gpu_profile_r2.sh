# round 2 ncu evidence IN THE BENCH WINDOW (BASELINE configs[2], 16384 games x 500 sims, warm-up 5 moves): the launch list of move 5 of the
# `value` pass and one `--set full` capture of the four kernels of a simulation step in the same move.  bench.py numbers printed under ncu are not bench values.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
# a simulation step = 6 kernels (select_expand, gc, expand_resume, tc_conv, tc_fc, backup); a move = 500 steps + ~8 launches; 5 warm-up moves
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 15200 --launch-count 1200 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 1 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/ncu_bench_r2.log 2>&1
tail -2 gpurun_out/launches_r2.csv | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tc_conv|k_tc_fc|k_select_expand|k_backup" --launch-skip 10100 --launch-count 8 \
    -o gpurun_out/prof_r2 python bench.py --steps 1 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/ncu_full_r2.log 2>&1
ls -la gpurun_out | tail -5
