#!/bin/bash
# round 2, call B: first run of the persistent kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bf16.py -x -q 2>&1 | tail -25 > gpurun_out/r02b_bf16.log
timeout 600 python -m pytest tests/test_gpu_cfg5_full.py -x -q -s 2>&1 | tail -30 > gpurun_out/r02b_cfg5.log
timeout 200 python bench.py --config cfg5 --precision bf16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_bf16.json 2> gpurun_out/r02b_bench_bf16.err
timeout 200 python bench.py --config cfg5 --precision bf16x3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_bf16x3.json 2> gpurun_out/r02b_bench_bf16x3.err
tail -5 gpurun_out/r02b_bf16.log; tail -5 gpurun_out/r02b_cfg5.log; tail -c 600 gpurun_out/r02b_bench_bf16.err; tail -c 300 gpurun_out/r02b_bench_bf16.json
