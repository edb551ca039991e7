cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
for h in 0 2560; do
  echo "=== gc_headroom $h, 16 moves, no timing"
  GC_HEADROOM=$h NO_TIMING=1 timeout 300 python scripts/exp_growth.py 16384 16384 500 16 net_tc 2>&1 | grep -E "^move" | cut -c1-200
done 2>&1 | tee gpurun_out/exp_gc.log
echo "=== timing, 4 moves" | tee -a gpurun_out/exp_gc.log
timeout 300 python scripts/exp_growth.py 16384 16384 500 4 net_tc 2>&1 | grep -E "^\{'select|k_select|S0_arrive" | cut -c1-700 | tee -a gpurun_out/exp_gc.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1_gc.json 2> gpurun_out/bench_r1_gc.err; python scripts/show_bench.py gpurun_out/bench_r1_gc.json; tail -3 gpurun_out/bench_r1_gc.err
