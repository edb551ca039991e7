cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_2iss.json 2> gpurun_out/bench_2iss.err; python scripts/show_bench.py gpurun_out/bench_2iss.json; tail -3 gpurun_out/bench_2iss.err
