# round 2, round-end validation on one B200 (path cache on by default): the whole -m gpu suite, smoke(), the driver's bench command for both arms.
# Every command under a hard (SIGKILL) limit.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
K="timeout -s KILL"
nvidia-smi --query-gpu=index,name --format=csv,noheader; nproc
$K 420 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
$K 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
$K 420 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_final2.json 2> gpurun_out/bench_r2_final2.err; tail -c 300 gpurun_out/bench_r2_final2.err; python scripts/show_bench.py gpurun_out/bench_r2_final2.json
$K 420 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/ref_r2_final2.json 2> gpurun_out/ref_r2_final2.err; tail -c 300 gpurun_out/ref_r2_final2.err; cut -c1-400 gpurun_out/ref_r2_final2.json
