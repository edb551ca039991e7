# round 2, round-end validation on one B200: the whole -m gpu suite, smoke(), the driver's bench command for both arms
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader; nproc
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; tail -c 300 gpurun_out/bench_r2_final.err; python scripts/show_bench.py gpurun_out/bench_r2_final.json
timeout 1500 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/ref_r2_final.json 2> gpurun_out/ref_r2_final.err; tail -c 300 gpurun_out/ref_r2_final.err; cut -c1-600 gpurun_out/ref_r2_final.json
