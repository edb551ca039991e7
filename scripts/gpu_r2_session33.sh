# round 2, GPU session 33: path cache, entry lines requested at kernel start / two rounds ahead (B200_PC_EARLY_PREFETCH=1, hint only) against the committed build, same box
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
K="timeout -s KILL"
for v in default pcpf; do
  echo "=== variant $v"
  if [ $v = default ]; then unset B200_TETRIS_LIB; else export B200_TETRIS_LIB=$GRAFT_REPO_ROOT/build/variants/lib_$v.so; fi
  if [ $v = pcpf ]; then $K 100 python -m pytest tests/test_gpu_engine.py -x -q -m gpu -k "path_cache" 2>&1 | tail -2; fi
  NO_TIMING=1 GC_HEADROOM=2560 $K 100 python scripts/exp_growth.py 16384 16384 500 8 net_tc 2>&1 | grep -E "^move  [4-7]" | cut -c1-130
  GC_HEADROOM=2560 $K 100 python scripts/exp_growth.py 16384 16384 500 7 net_tc 2>&1 | grep -E "^move  [5-6]" | cut -c1-330
done 2>&1 | tee gpurun_out/exp_path_cache_r2g.txt
