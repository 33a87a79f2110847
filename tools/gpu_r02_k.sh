#!/bin/bash
# round 2, second half: 2-GPU evidence with the pipelined backward pass (NCCL gradient check, weak and strong scaling) + 8192 series on one GPU
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --check > gpurun_out/r02k_check2.json 2> gpurun_out/r02k_check2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02k_bench_2gpu.json 2> gpurun_out/r02k_bench_2gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --global-batch 8192 --no-cpu-baseline > gpurun_out/r02k_bench_2gpu_strong.json 2> gpurun_out/r02k_bench_2gpu_strong.err
timeout 400 python bench.py --steps 5 --warmup 3 --global-batch 8192 --no-cpu-baseline --no-alt-mode > gpurun_out/r02k_bench_1gpu_strong.json 2> gpurun_out/r02k_bench_1gpu_strong.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-alt-mode > gpurun_out/r02k_bench_1gpu.json 2> gpurun_out/r02k_bench_1gpu.err
for f in check2 bench_2gpu bench_2gpu_strong bench_1gpu_strong bench_1gpu; do echo "== $f"; tail -c 600 gpurun_out/r02k_$f.json | cut -c1-600; tail -c 300 gpurun_out/r02k_$f.err; done
