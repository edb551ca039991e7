# round 2, GPU session 3: conv front-end restructure + trainer + wire formats: parity suite, phase split, the driver's bench command
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) | tee gpurun_out/pytest_gpu_r2c.log
timeout 300 python scripts/exp_growth.py 16384 16384 500 6 net_tc 2>&1 | grep -E "^move  [15]|^\{'select|k_select|S0_arrive" | cut -c1-700 | tee gpurun_out/exp_r2c.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; tail -c 400 gpurun_out/bench_r2c.err; python scripts/show_bench.py gpurun_out/bench_r2c.json
