# round 2, GPU session 17: is the trace length of a game predictable (deep lane)?  phase split of the distributional configuration
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export B200_TETRIS_LIB=$GRAFT_REPO_ROOT/build/variants/lib_nopair.so
timeout 600 python scripts/exp_trace_corr.py 16384 16384 500 9 2>&1 | tail -8 | tee gpurun_out/exp_trace_corr_r2.txt
timeout 600 python scripts/exp_dist_phases.py 2048 32768 1500 2 2>&1 | tail -6 | tee gpurun_out/exp_dist_phases_r2.txt
timeout 600 python scripts/exp_dist_phases.py 16384 16384 500 2 2>&1 | tail -6 | tee -a gpurun_out/exp_dist_phases_r2.txt
