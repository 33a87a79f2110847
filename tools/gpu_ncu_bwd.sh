#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:persist_ -s 2 -c 2 -o gpurun_out/r02l_persist_bf16x3 -f python tools/one_step.py bf16x3 2 > gpurun_out/r02l_ncu_bf16x3.log 2>&1
tail -3 gpurun_out/r02l_ncu_bf16x3.log
