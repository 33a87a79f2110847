#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
cat > /tmp/one.py <<'PY'
import os, sys
R = os.getcwd()
sys.path[:0] = [R, os.path.join(R, "online-neural-cdes_b200")]
import torch, bench
import torchcde_b200 as tc
from ncde_b200 import OriginalVectorField
prec = sys.argv[1]
cfg = bench.CFG
x, _, _ = bench.synth_batch(1024, 3)
c = tc.linear_interpolation_coeffs(x.cuda(), rectilinear=0)
torch.manual_seed(0)
f = OriginalVectorField(cfg["C"], cfg["H"], cfg["HH"], cfg["n_layers"]).cuda()
z0 = (torch.randn(1024, 128, device="cuda") * 0.5).requires_grad_(True)
X = tc.LinearInterpolation(c)
for _ in range(2):
    out = tc.cdeint(X, f, z0, X.grid_points, adjoint=False, method="rk4", options={"step_size": 1, "precision": prec})
    out.sum().backward()
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:persist_ -s 2 -c 2 -o gpurun_out/r02d_persist_bf16 -f python /tmp/one.py bf16 > gpurun_out/r02d_ncu.log 2>&1
tail -5 gpurun_out/r02d_ncu.log
