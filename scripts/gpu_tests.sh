set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 600 python scripts/exp_growth.py 1024 8192 500 12 net 2>&1 | tail -14
timeout 600 python scripts/exp_growth.py 16384 8192 500 9 net 2>&1 | tail -11
