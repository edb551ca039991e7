# round 2, GPU session 15: conv3 pair mode with the re-dealt conv2 store: parity + A/B (E2|E3|E1, E3|E2|E1, 8-wide conv3)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for v in default e3first nopair; do
  echo "=== variant $v"
  if [ $v = default ]; then unset B200_TETRIS_LIB; else export B200_TETRIS_LIB=$GRAFT_REPO_ROOT/build/variants/lib_$v.so; fi
  timeout 300 python scripts/dbg_conv_pair.py 2>&1 | grep -v "n_bad 0" | tail -5
  GC_HEADROOM=2560 timeout 600 python scripts/exp_growth.py 16384 16384 500 8 net_tc 2>&1 | grep -E "^move  [67]|wait_c2" | cut -c1-520
done 2>&1 | tee gpurun_out/exp_variants_r2i.txt
