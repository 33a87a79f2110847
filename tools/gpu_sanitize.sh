#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for m in bf16x3 bf16; do
  timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_case.py $m > gpurun_out/sanitize_$m.log 2>&1
  echo "== $m: $(grep -c '^ok' gpurun_out/sanitize_$m.log) cases ok; $(grep 'ERROR SUMMARY' gpurun_out/sanitize_$m.log | tail -1)"
  grep -E "Invalid|Unknown Error|out of bounds|misaligned" gpurun_out/sanitize_$m.log | sort | uniq -c | head -5
done
