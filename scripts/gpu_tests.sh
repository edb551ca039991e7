set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 900 python bench.py --workload dist --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_dist.json 2> gpurun_out/bench_r1_dist.err; python scripts/show_bench.py gpurun_out/bench_r1_dist.json; tail -3 gpurun_out/bench_r1_dist.err
