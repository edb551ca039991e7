# round 2, GPU session 6: two-set conv pipeline (vs the one-set form), plain row loads; network parity tests first
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_valuenet.py tests/test_gpu_engine.py tests/test_gpu_bench_config.py -x -q 2>&1 | tail -6 ) | tee gpurun_out/pytest_gpu_r2f.log
for v in default sets1 rowplain; do
  echo "=== variant $v"
  if [ $v = default ]; then unset B200_TETRIS_LIB; else export B200_TETRIS_LIB=$GRAFT_REPO_ROOT/build/variants/lib_$v.so; fi
  GC_HEADROOM=2560 timeout 300 python scripts/exp_growth.py 16384 16384 500 10 net_tc 2>&1 | grep -E "^move  [2579]|^\{'select|S0_arrive" | cut -c1-500
done 2>&1 | tee gpurun_out/exp_variants_r2f.log
