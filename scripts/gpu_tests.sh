set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 600 python scripts/exp_growth.py 16384 16384 500 4 net_tc 2>&1 | tail -7
