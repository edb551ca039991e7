# round 2, GPU session 16: ncu --set full of k_tc_conv in the bench window, conv3 pair mode against the 8-wide conv3
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for v in default nopair; do
  if [ $v = default ]; then unset B200_TETRIS_LIB; else export B200_TETRIS_LIB=$GRAFT_REPO_ROOT/build/variants/lib_$v.so; fi
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tc_conv" --launch-skip 2600 --launch-count 1 \
      -o gpurun_out/prof_conv_$v python bench.py --steps 1 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/ncu_conv_$v.log 2>&1
  tail -2 gpurun_out/ncu_conv_$v.log | cut -c1-300
done
ls -la gpurun_out | tail -4
