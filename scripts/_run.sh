cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_valuenet.py -x -q 2>&1 | tail -3
timeout 300 python scripts/exp_growth.py 16384 16384 200 1 net_tc 2>&1 | tail -3 | cut -c1-900
