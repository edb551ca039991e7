# round 2, GPU session 10 (1 GPU): replay-policy tests + smoke()
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_replay_policy.py tests/test_gpu_engine.py -q 2>&1 | tail -12 ) | tee gpurun_out/pytest_gpu_r2h.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
