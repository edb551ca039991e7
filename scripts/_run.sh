cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/exp_growth.py 16384 16384 500 2 synthetic 2>&1 | grep -E "^move|select_expand" | cut -c1-330
