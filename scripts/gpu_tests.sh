set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1_tc.json 2> gpurun_out/bench_r1_tc.err; tail -c 3500 gpurun_out/bench_r1_tc.json; tail -5 gpurun_out/bench_r1_tc.err
