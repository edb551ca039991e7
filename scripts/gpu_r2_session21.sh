# round 2, GPU session 21: rolling L2 prefetch along the previous trace inside the walk (B200_ROLL_PREFETCH=1) against the committed walk
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for v in default roll; do
  echo "=== variant $v"
  if [ $v = default ]; then unset B200_TETRIS_LIB; else export B200_TETRIS_LIB=$GRAFT_REPO_ROOT/build/variants/lib_$v.so; fi
  GC_HEADROOM=2560 timeout 600 python scripts/exp_growth.py 16384 16384 500 8 net_tc 2>&1 | grep -E "^move  [3-7]|k_select_expand sampled" | cut -c1-330
  NO_TIMING=1 GC_HEADROOM=2560 timeout 600 python scripts/exp_growth.py 16384 16384 500 8 net_tc 2>&1 | grep -E "^move  [5-7]" | cut -c1-130
done 2>&1 | tee gpurun_out/exp_variants_r2k.txt
export B200_TETRIS_LIB=$GRAFT_REPO_ROOT/build/variants/lib_roll.so
timeout 900 python -m pytest tests/test_gpu_bench_config.py tests/test_gpu_engine.py -x -q -m gpu 2>&1 | tail -3
