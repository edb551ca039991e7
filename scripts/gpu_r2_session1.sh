# round 2, GPU session 1: parity suite (incl. the bench-config shadow test), the restructured bench line, the reference arm on the
# box's cores, and the A/B builds of k_select_expand (per-level clock split, one game per warp, no warm loads)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc
( timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -25 ) | tee gpurun_out/pytest_gpu_r2a.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -c 600 gpurun_out/bench_r2a.err; python scripts/show_bench.py gpurun_out/bench_r2a.json
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/ref_r2a.json 2> gpurun_out/ref_r2a.err ) 2>&1 | tail -4; tail -c 300 gpurun_out/ref_r2a.err; cut -c1-300 gpurun_out/ref_r2a.json
for v in prof onegame nowarm; do
  echo "=== variant $v"
  export B200_TETRIS_LIB=$GRAFT_REPO_ROOT/build/variants/lib_$v.so
  timeout 300 python scripts/exp_growth.py 16384 16384 500 6 net_tc 2>&1 | grep -E "^move|^\{'select|k_select|per level|S0_arrive" | cut -c1-600
done 2>&1 | tee gpurun_out/exp_variants_r2a.log
