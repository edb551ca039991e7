# round 2, GPU session 29: where the walk's time goes with the path cache on (B200_SELECT_PROF build: phase 1 clocks / rounds / levels served, per-level split of the uncached levels)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
K="timeout -s KILL"
export B200_TETRIS_LIB=$GRAFT_REPO_ROOT/build/variants/lib_selprof.so
GC_HEADROOM=2560 $K 150 python scripts/exp_growth.py 16384 16384 500 8 net_tc 2>&1 | grep -E "^move  [4-7]|k_select_expand sampled|per level|path cache phase" | cut -c1-330 | tee gpurun_out/exp_path_cache_prof_r2.txt
