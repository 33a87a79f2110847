#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --check > gpurun_out/r02h_check2.json 2> gpurun_out/r02h_check2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02h_bench_2gpu.json 2> gpurun_out/r02h_bench_2gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --global-batch 8192 > gpurun_out/r02h_bench_2gpu_strong.json 2> gpurun_out/r02h_bench_2gpu_strong.err
tail -2 gpurun_out/r02h_check2.json; tail -c 400 gpurun_out/r02h_bench_2gpu.json; tail -c 300 gpurun_out/r02h_check2.err
