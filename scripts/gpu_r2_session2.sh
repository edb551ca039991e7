# round 2, GPU session 2: lockstep walk (GrpW) + WARM_SELECT off + mbarrier fast path: parity suite, then A/B builds
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) | tee gpurun_out/pytest_gpu_r2b.log
for v in default prof twogame onegame nofast nohot nowarmx; do
  echo "=== variant $v"
  if [ $v = default ]; then unset B200_TETRIS_LIB; else export B200_TETRIS_LIB=$GRAFT_REPO_ROOT/build/variants/lib_$v.so; fi
  timeout 300 python scripts/exp_growth.py 16384 16384 500 6 net_tc 2>&1 | grep -E "^move  [15]|^\{'select|k_select|per level|S0_arrive" | cut -c1-600
done 2>&1 | tee gpurun_out/exp_variants_r2b.log
