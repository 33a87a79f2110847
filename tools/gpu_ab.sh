#!/bin/bash
# A/B on ONE box: tools/micro/libncde_old.so against the in-tree build (bench line of both tensor-core modes, twice each, interleaved)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
LIB=online-neural-cdes_b200/torchcde_b200/libncde_b200.so
cp $LIB /tmp/new.so
timeout 600 python -m pytest tests/test_gpu_cfg5_full.py tests/test_gpu_bf16.py tests/test_reverse_perturb.py -m gpu -x -q 2>&1 | tail -2
for rep in 1 2; do
for v in old new; do
  if [ $v = old ]; then cp tools/micro/libncde_old.so $LIB; else cp /tmp/new.so $LIB; fi
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ab_bench.json 2> gpurun_out/ab_bench.err
  python - $v <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/ab_bench.json").read().strip().splitlines()[-1])
    om = d.get("other_modes") or {}
    print(sys.argv[1], "x3 ms/step %.2f" % d["ms_per_step"], {k: round(v["ms_per_step"], 2) for k, v in d.get("kernel_ms", {}).items() if k.startswith("solve")}, "| bf16 %.2f" % list(om.values())[0]["ms_per_step"] if om else "")
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/ab_bench.err").read()[-600:])
PY
done
done
cp /tmp/new.so $LIB
