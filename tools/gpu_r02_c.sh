#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 120 python tools/ps_trace.py bf16 2> gpurun_out/r02c_trace_bf16.txt > /dev/null
timeout 120 python tools/ps_trace.py bf16x3 2> gpurun_out/r02c_trace_bf16x3.txt > /dev/null
python tools/trace_stats.py gpurun_out/r02c_trace_bf16.txt; python tools/trace_stats.py gpurun_out/r02c_trace_bf16x3.txt; timeout 300 python -m pytest tests/test_gpu_bf16.py -x -q 2>&1 | tail -3
