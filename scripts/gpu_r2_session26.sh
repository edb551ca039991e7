# round 2, GPU session 26: the path cache (b200_set_path_cache) — parity first (its own tests, then the engine / bench-config / agent suites with the
# cache switched on everywhere through B200_PATH_CACHE=1), then the per-kernel times of moves 3-7 with the cache off and on
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py -x -q -m gpu -k "path_cache" 2>&1 | tail -15
B200_PATH_CACHE=1 timeout 1500 python -m pytest tests/test_gpu_engine.py tests/test_gpu_bench_config.py tests/test_gpu_agents.py -x -q -m gpu 2>&1 | tail -6
for pc in 0 1; do
  echo "=== B200_PATH_CACHE=$pc"
  B200_PATH_CACHE=$pc GC_HEADROOM=2560 timeout 600 python scripts/exp_growth.py 16384 16384 500 8 net_tc 2>&1 | grep -E "^move  [2-7]|k_select_expand sampled" | cut -c1-330
  B200_PATH_CACHE=$pc NO_TIMING=1 GC_HEADROOM=2560 timeout 600 python scripts/exp_growth.py 16384 16384 500 8 net_tc 2>&1 | grep -E "^move  [4-7]" | cut -c1-130
done 2>&1 | tee gpurun_out/exp_path_cache_r2c.txt
