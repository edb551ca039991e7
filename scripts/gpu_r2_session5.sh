# round 2, GPU session 5: principal-variation prefetch variants (per-level clock split), bench-window regime
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for v in pv0 pv1 pv2 pv3 default; do
  echo "=== variant $v"
  if [ $v = default ]; then unset B200_TETRIS_LIB; else export B200_TETRIS_LIB=$GRAFT_REPO_ROOT/build/variants/lib_$v.so; fi
  GC_HEADROOM=2560 timeout 300 python scripts/exp_growth.py 16384 16384 500 10 net_tc 2>&1 | grep -E "^move  [2579]|^\{'select|k_select|per level" | cut -c1-420
done 2>&1 | tee gpurun_out/exp_variants_r2e.log
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) | tee gpurun_out/pytest_gpu_r2e.log
