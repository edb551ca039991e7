# round 2, GPU session 13: locate the conv3-pair defect (variants: padded pair rows / unpadded / 8-wide conv3)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for v in default psbo128 nopair; do
  echo "=== variant $v"
  if [ $v = default ]; then unset B200_TETRIS_LIB; else export B200_TETRIS_LIB=$GRAFT_REPO_ROOT/build/variants/lib_$v.so; fi
  timeout 300 python scripts/dbg_conv_pair.py 2>&1 | tail -12
done
unset B200_TETRIS_LIB
timeout 900 python scripts/exp_two_engines.py 16384 16384 500 9 1,2,4 2>&1 | tail -5
