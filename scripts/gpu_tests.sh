set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_valuenet.py -x -q -m gpu 2>&1 | tail -12
timeout 600 python scripts/exp_growth.py 16384 16384 500 4 net_tc 2>&1 | tail -6
