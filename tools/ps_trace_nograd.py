"""Debug: forward-only (no saved records) cfg-5-shaped solve with NCDE_PS_TRACE=1: the hidden role's chain without its record stores."""
import os, sys
os.environ["NCDE_PS_TRACE"] = "1"
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "online-neural-cdes_b200")]
import torch
import bench
import torchcde_b200 as tc
from ncde_b200 import OriginalVectorField
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
cfg = bench.CFG
x, _, _ = bench.synth_batch(1024, 3)
c = tc.linear_interpolation_coeffs(x.cuda(), rectilinear=0)
torch.manual_seed(0)
f = OriginalVectorField(cfg["C"], cfg["H"], cfg["HH"], cfg["n_layers"]).cuda()
z0 = torch.randn(1024, 128, device="cuda") * 0.5
X = tc.LinearInterpolation(c)
with torch.no_grad():
    for _ in range(2):
        out = tc.cdeint(X, f, z0, X.grid_points, adjoint=False, method="rk4", options={"step_size": 1, "precision": prec})
torch.cuda.synchronize()
