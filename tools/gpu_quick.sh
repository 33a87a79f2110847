#!/bin/bash
# quick loop for kernel work: the persistent-path parity tests, then the bench line of both tensor-core modes (no CPU baseline)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cfg5_full.py tests/test_gpu_bf16.py tests/test_reverse_perturb.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/quick_bench.json").read().strip().splitlines()[-1])
    print("MODE", d["dtype"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    print({k: round(v["ms_per_step"], 2) for k, v in d.get("kernel_ms", {}).items()})
    for k, v in (d.get("other_modes") or {}).items():
        print("OTHER", k, v)
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/quick_bench.err").read()[-1500:])
PY
