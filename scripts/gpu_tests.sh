set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1_occ.json 2> gpurun_out/bench_r1_occ.err; python scripts/show_bench.py gpurun_out/bench_r1_occ.json; tail -3 gpurun_out/bench_r1_occ.err
