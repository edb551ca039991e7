# round 2, GPU session 11 (2 GPUs): the driver's multi-GPU launch of both arms, short
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 3 --warmup 5 > gpurun_out/bench_r2_2gpu.json 2> gpurun_out/bench_r2_2gpu.err; tail -c 600 gpurun_out/bench_r2_2gpu.err; python scripts/show_bench.py gpurun_out/bench_r2_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/ref_r2_2gpu.json 2> /dev/null; cut -c1-200 gpurun_out/ref_r2_2gpu.json
GC_HEADROOM=2560 NO_TIMING=1 timeout 300 python scripts/exp_growth.py 16384 16384 500 12 net_tc 2>&1 | grep -E "^move 1[01]|trace length" | cut -c1-250
