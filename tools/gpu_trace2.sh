#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for m in bf16x3 bf16; do
  timeout 200 python tools/ps_trace_nograd.py $m 2> gpurun_out/trace_ng_$m.txt > /dev/null
  echo "== nograd $m"; python tools/trace_stats.py gpurun_out/trace_ng_$m.txt 2>&1 | grep -v "t=[1-7]:" | cut -c1-330
  timeout 200 python tools/ps_trace.py $m 2> gpurun_out/trace_g_$m.txt > /dev/null
  echo "== grad $m"; python tools/trace_stats.py gpurun_out/trace_g_$m.txt 2>&1 | grep "^fwd\|fwd.*t=0\|q=.*t=0:.*z_in" | cut -c1-330
done
