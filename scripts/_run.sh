cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 60 python -m pytest tests/test_gpu_engine.py -k "valuesim_and_vanilla" -q ) 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_modes.log
