cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
for v in base we we_ws we_st default we_l2_8 we_l2_32 we_l2_b12 we_l2_b16; do
  echo "=== variant $v"
  if [ $v = default ]; then unset B200_TETRIS_LIB; else export B200_TETRIS_LIB=$GRAFT_REPO_ROOT/build/variants/lib_$v.so; fi
  timeout 300 python scripts/exp_growth.py 16384 16384 500 4 net_tc 2>&1 | grep -E "^move|^\{'select|k_select" | cut -c1-400
done 2>&1 | tee gpurun_out/exp_variants.log
unset B200_TETRIS_LIB
echo "=== default, no timing (graph replay)"
NO_TIMING=1 timeout 300 python scripts/exp_growth.py 16384 16384 500 4 net_tc 2>&1 | grep -E "^move" | cut -c1-200 | tee -a gpurun_out/exp_variants.log
echo "=== default lib, M=4096 (footprint sensitivity)" | tee -a gpurun_out/exp_variants.log
timeout 300 python scripts/exp_growth.py 16384 4096 500 4 net_tc 2>&1 | grep -E "^move|^\{'select|k_select" | cut -c1-400 | tee -a gpurun_out/exp_variants.log
