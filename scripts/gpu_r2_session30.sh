# round 2, GPU session 30: one ncu --set full capture (with source) of k_select_expand and k_backup with the path cache on, move 5 of the bench configuration
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
K="timeout -s KILL"
GC_HEADROOM=2560 $K 400 ncu --set full --clock-control none --import-source on -k regex:"k_select_expand|k_backup" --launch-skip 5200 --launch-count 2 -f -o gpurun_out/prof_pc_r2 \
    python scripts/exp_growth.py 16384 16384 500 6 net_tc > gpurun_out/ncu_pc_r2.log 2>&1
tail -3 gpurun_out/ncu_pc_r2.log | cut -c1-200; ls -la gpurun_out/ | tail -4
