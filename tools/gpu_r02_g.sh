#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02g_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02g_smoke.log 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r02g_bench_default.json 2> gpurun_out/r02g_bench_default.err
timeout 400 python bench.py --steps 3 --warmup 3 --global-batch 8192 --no-cpu-baseline > gpurun_out/r02g_bench_strong1.json 2> gpurun_out/r02g_bench_strong1.err
tail -3 gpurun_out/r02g_tests.log; tail -3 gpurun_out/r02g_smoke.log; tail -c 300 gpurun_out/r02g_bench_default.err; tail -c 300 gpurun_out/r02g_bench_strong1.err
