# round 2, GPU session 4: conv issuer order (conv1 first), trace records for the backup, set_games without a second collection;
# per-move phases over the bench window (collections, tree drops) with the longest trace of every move
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 ) | tee gpurun_out/pytest_gpu_r2d.log
GC_HEADROOM=2560 timeout 400 python scripts/exp_growth.py 16384 16384 500 14 net_tc 2>&1 | grep -E "^move|^\{'select|k_select|S0_arrive" | cut -c1-700 | tee gpurun_out/exp_r2d.log
timeout 900 python bench.py --steps 10 --warmup 5 --no-secondary > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; tail -c 400 gpurun_out/bench_r2d.err; python scripts/show_bench.py gpurun_out/bench_r2d.json
