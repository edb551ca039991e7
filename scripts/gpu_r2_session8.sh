# round 2, GPU session 8 (1 GPU): full parity suite incl. the replay-policy tests
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) | tee gpurun_out/pytest_gpu_r2g.log
