# round 2, last GPU session: smoke() and a short bench line on the exact round-end build (pc_eval v6)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
K="timeout -s KILL"
$K 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
$K 200 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/bench_r2_final3.json 2> gpurun_out/bench_r2_final3.err; tail -c 200 gpurun_out/bench_r2_final3.err; python scripts/show_bench.py gpurun_out/bench_r2_final3.json 2>/dev/null | head -5 | cut -c1-420
