# round 2, GPU session 22 (re-created container): random-access probe of the memory system + a short bench of the rebuilt library
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 300 ./scripts/probe/mem_probe 2>&1 | tee gpurun_out/mem_probe_r2.txt
timeout 600 python bench.py --steps 5 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/bench_r2s22.json 2> gpurun_out/bench_r2s22.err; tail -c 300 gpurun_out/bench_r2s22.err; python scripts/show_bench.py gpurun_out/bench_r2s22.json
