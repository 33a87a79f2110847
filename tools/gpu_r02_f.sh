#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_cfg5_full.py -q 2>&1 | tail -4
for hg in 0 1; do
  for p in bf16 bf16x3; do
    NCDE_TC_HG=$hg timeout 200 python bench.py --config cfg5 --precision $p --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('HG=$hg', '$p', round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['kernel_ms'].items()})"
  done
done
