# round 2, GPU session 20: scale factors folded into the bias fma of the conv epilogues (bit-identical): parity + timing
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_valuenet.py tests/test_gpu_dist.py tests/test_gpu_bench_config.py -x -q -m gpu 2>&1 | tail -4
GC_HEADROOM=2560 timeout 600 python scripts/exp_growth.py 16384 16384 500 8 net_tc 2>&1 | grep -E "^move  [67]|wait_c2" | cut -c1-520
