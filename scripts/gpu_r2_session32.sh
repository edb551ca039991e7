# round 2, GPU session 32: path cache v3 (one lane per cached level).  Every command under a hard (SIGKILL) time limit and chained: session 26
# hung in the first test (a cached terminal leaf returned the wrong node, the arena got corrupted, a hash-table probe never ended) and a SIGTERM
# does not reach a process blocked in cudaDeviceSynchronize.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
K="timeout -s KILL"
$K 150 python -m pytest tests/test_gpu_engine.py -x -q -m gpu -k "path_cache" 2>&1 | tail -5; test ${PIPESTATUS[0]} -eq 0 || { echo "path-cache tests failed: stop"; exit 1; }
B200_PATH_CACHE=1 $K 300 python -m pytest tests/test_gpu_engine.py tests/test_gpu_bench_config.py tests/test_gpu_agents.py -x -q -m gpu 2>&1 | tail -5; test ${PIPESTATUS[0]} -eq 0 || { echo "suite with the cache on failed: stop"; exit 1; }
for pc in 1; do
  echo "=== B200_PATH_CACHE=$pc"
  B200_PATH_CACHE=$pc GC_HEADROOM=2560 $K 120 python scripts/exp_growth.py 16384 16384 500 8 net_tc 2>&1 | grep -E "^move  [2-7]|k_select_expand sampled" | cut -c1-330
  B200_PATH_CACHE=$pc NO_TIMING=1 GC_HEADROOM=2560 $K 120 python scripts/exp_growth.py 16384 16384 500 8 net_tc 2>&1 | grep -E "^move  [4-7]" | cut -c1-130
done 2>&1 | tee gpurun_out/exp_path_cache_r2f.txt
