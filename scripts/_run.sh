cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python scripts/exp_growth.py 16384 16384 200 1 net_tc 2>&1 | tail -4 | cut -c1-900
