# round 2, GPU session 18: distributional network on tcgen05 (k_tdc_conv / k_tdc_fc): parity, then the phase split; mbarrier suspend-hint A/B
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_valuenet.py -x -q -m gpu 2>&1 | tail -15
sed -i 's/eval_kind=.net., dist_weights/eval_kind=os.environ.get("DIST_EVAL", "net"), dist_weights/; s/^import sys, time/import os, sys, time/' scripts/exp_dist_phases.py
DIST_EVAL=net_tc timeout 600 python scripts/exp_dist_phases.py 2048 32768 1500 2 2>&1 | tail -6
DIST_EVAL=net_tc timeout 600 python scripts/exp_dist_phases.py 16384 16384 500 2 2>&1 | tail -6
for v in default suspend; do
  echo "=== variant $v"
  if [ $v = default ]; then unset B200_TETRIS_LIB; else export B200_TETRIS_LIB=$GRAFT_REPO_ROOT/build/variants/lib_$v.so; fi
  GC_HEADROOM=2560 timeout 600 python scripts/exp_growth.py 16384 16384 500 8 net_tc 2>&1 | grep -E "^move  [67]|wait_c2" | cut -c1-520
done 2>&1 | tee gpurun_out/exp_variants_r2j.txt
