# round 2, GPU session 12: conv3 on the transposed pair grid — network parity, search parity at the bench configuration, conv clk per board, bench
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m pytest tests/test_gpu_valuenet.py tests/test_gpu_agents.py -x -q -m gpu 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_bench_config.py tests/test_gpu_engine.py tests/test_gpu_replay_policy.py -x -q -m gpu 2>&1 | tail -5
GC_HEADROOM=2560 timeout 600 python scripts/exp_growth.py 16384 16384 500 8 net_tc 2>&1 | grep -E "^move [67]|conv|per board|trace length" | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 5 > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err; tail -c 400 gpurun_out/bench_r2h.err; python scripts/show_bench.py gpurun_out/bench_r2h.json
