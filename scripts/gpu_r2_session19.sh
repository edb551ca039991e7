# round 2, GPU session 19: deep lane: exactness (every game, against an engine without it; bench-config shadow test), then A/B over the lane size
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_bench_config.py -x -q -m gpu 2>&1 | tail -8
for lane in 0 82 164 328 656; do
  echo "=== deep lane $lane"
  DEEP_LANE=$lane NO_TIMING=1 GC_HEADROOM=2560 timeout 600 python scripts/exp_growth.py 16384 16384 500 10 net_tc 2>&1 | grep -E "^move  [5-9]" | cut -c1-130
done 2>&1 | tee gpurun_out/exp_deep_lane_r2.txt
