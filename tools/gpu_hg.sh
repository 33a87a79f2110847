#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for hg in 1 2; do
NCDE_TC_HG=$hg timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/hg_bench.json 2> gpurun_out/hg_bench.err
python - $hg <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/hg_bench.json").read().strip().splitlines()[-1])
    print("HG", sys.argv[1], d["dtype"][:12], "ms/step", round(d["ms_per_step"], 2), {k: round(v["ms_per_step"], 2) for k, v in d.get("kernel_ms", {}).items()})
    for k, v in (d.get("other_modes") or {}).items():
        print("   OTHER", k, round(v["ms_per_step"], 2), v.get("kernel_ms"))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/hg_bench.err").read()[-800:])
PY
done
