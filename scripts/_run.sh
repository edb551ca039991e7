cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q ) 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
