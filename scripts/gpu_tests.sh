set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1_fp16x2.json 2> gpurun_out/bench_r1_fp16x2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_fp16x2.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['cpu_baseline']['value'], d['phases_ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline_select_backup']['achieved'], d['clocks'])"; tail -3 gpurun_out/bench_r1_fp16x2.err
