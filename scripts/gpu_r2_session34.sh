# round 2, GPU session 34: pc_eval with the picked slot's terms computed once and a branch-free second pass: parity (the whole -m gpu suite), then timing
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
K="timeout -s KILL"
$K 200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3; test ${PIPESTATUS[0]} -eq 0 || { echo "gpu suite failed: stop"; exit 1; }
{ NO_TIMING=1 GC_HEADROOM=2560 $K 100 python scripts/exp_growth.py 16384 16384 500 8 net_tc 2>&1 | grep -E "^move  [4-7]" | cut -c1-130
  GC_HEADROOM=2560 $K 100 python scripts/exp_growth.py 16384 16384 500 7 net_tc 2>&1 | grep -E "^move  [5-6]|k_select_expand sampled" | cut -c1-330; } 2>&1 | tee gpurun_out/exp_path_cache_r2h.txt
