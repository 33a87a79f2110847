#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
LIB=online-neural-cdes_b200/torchcde_b200/libncde_b200.so
cp $LIB /tmp/orig.so
for v in P; do
  cp tools/micro/libncde_dbg_$v.so $LIB
  echo "== variant $v"
  timeout 120 python tools/sanitize_case.py bf16x3 2>&1 | grep -v "^  File\|^    " | grep -v "^  File\|^    " | cut -c1-200 | sort | uniq -c | sort -rn | head -30
done
cp /tmp/orig.so $LIB
