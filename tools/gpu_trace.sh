#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for m in ${MODES:-bf16x3 bf16}; do
  for g in ${TG:-0}; do
    NCDE_PS_TRACE_G=$g timeout 200 python tools/ps_trace.py $m 2> gpurun_out/trace_${m}_g$g.txt > /dev/null
    echo "== $m g=$g"; python tools/trace_stats.py gpurun_out/trace_${m}_g$g.txt 2>&1 | cut -c1-420; python tools/trace_all.py gpurun_out/trace_${m}_g$g.txt
  done
done
