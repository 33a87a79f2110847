#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/final_gputests.txt; cat gpurun_out/final_gputests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 >> gpurun_out/final_gputests.txt; tail -1 gpurun_out/final_gputests.txt
timeout 400 python bench.py --steps 10 --warmup 3 --cpu-baseline-seconds 10 > gpurun_out/final_bench_cfg5.json 2> gpurun_out/final_bench_cfg5.err
tail -c 300 gpurun_out/final_bench_cfg5.json
