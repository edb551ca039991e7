# default GPU-box job (gpurun -- 'bash scripts/_run.sh'): parity suite + the bench line
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q ) 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; python scripts/show_bench.py gpurun_out/bench.json; tail -3 gpurun_out/bench.err
