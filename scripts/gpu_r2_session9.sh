# round 2, GPU session 9 (1 GPU): full parity suite (replay policies, 1.5 M tables, pooled GC scratch), arena-size comparison, the bench line with every BASELINE configuration
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) | tee gpurun_out/pytest_gpu_r2g.log
for M in 16384 24576; do
  echo "=== max_nodes $M"
  GC_HEADROOM=$((M*5/32)) NO_TIMING=1 timeout 400 python scripts/exp_growth.py 16384 $M 500 22 net_tc 2>&1 | grep -E "^move" | cut -c1-200
done 2>&1 | tee gpurun_out/exp_arena_r2g.log
timeout 900 python bench.py --steps 5 --warmup 5 > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; tail -c 300 gpurun_out/bench_r2g.err; python scripts/show_bench.py gpurun_out/bench_r2g.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2g.json"))
for k in d:
    if k.startswith("also_"):
        print(k, {a: (round(b, 2) if isinstance(b, float) else b) for a, b in d[k].items() if a != "workload"})
PY
