# round 2, GPU session 23: the explicit-stream entry points on the GPU, and the default bench line with whole-episode Vanilla timing
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -x -q -m gpu -k "callers_stream or graph_replay" 2>&1 | tail -4
timeout 900 python bench.py --steps 5 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2s23.json 2> gpurun_out/bench_r2s23.err; tail -c 300 gpurun_out/bench_r2s23.err; python scripts/show_bench.py gpurun_out/bench_r2s23.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2s23.json"))
print(json.dumps(d.get("also_configs1_vanilla"), indent=1))
print(json.dumps(d.get("roofline_select_backup", {}).get("random_access"), indent=1))
PY
