#!/bin/bash
# timing-only A/B of debug variants (results of the variant are WRONG: no tests)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
LIB=online-neural-cdes_b200/torchcde_b200/libncde_b200.so
cp $LIB /tmp/new.so
for v in old $VARIANTS old $VARIANTS; do
  cp tools/micro/libncde_$v.so $LIB
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-alt-mode > gpurun_out/ab_bench.json 2> gpurun_out/ab_bench.err
  python - $v <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/ab_bench.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "x3 ms/step %.2f" % d["ms_per_step"], {k: round(v["ms_per_step"], 2) for k, v in d.get("kernel_ms", {}).items() if k.startswith("solve")})
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/ab_bench.err").read()[-600:])
PY
done
cp /tmp/new.so $LIB
