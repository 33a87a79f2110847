"""Small bf16 + fp32 fwd/bwd solves, dopri5 + adjoint: the workload compute-sanitizer is run on."""
import os, sys, copy
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "online-neural-cdes_b200")]
import torch
from oracle import cde_oracle as O
import torchcde_b200 as tc

torch.manual_seed(0)
for (B, L, C, H, HH, n) in [(130, 3, 100, 128, 128, 3), (70, 3, 5, 16, 32, 2)]:
    x = torch.randn(B, L, C); x[..., 0] = torch.arange(L, dtype=torch.float32)
    func = O.SharedMLPField(C, H, HH, n).cuda()
    coeffs = tc.linear_interpolation_coeffs(x.cuda(), rectilinear=0)
    X = tc.LinearInterpolation(coeffs)
    for prec in (sys.argv[1:2] or ["bf16", "fp32"]) if len(sys.argv) > 1 else ("bf16", "fp32"):
        z0 = (torch.randn(B, H) * 0.5).cuda().requires_grad_(True)
        out = tc.cdeint(X, func, z0, X.grid_points, adjoint=False, method="rk4", options={"step_size": 1, "precision": prec})
        out.sum().backward()
        torch.cuda.synchronize()
        print("ok", B, C, H, prec, float(out.abs().max()))
if len(sys.argv) > 1:
    sys.exit(0)
x = torch.randn(6, 6, 4); x[..., 0] = torch.arange(6, dtype=torch.float32)
Xc = tc.NaturalCubicSpline(tc.natural_cubic_coeffs(x.cuda()))
func = O.SharedMLPField(4, 8, 8, 2).cuda()
z0 = (torch.randn(6, 8) * 0.5).cuda().requires_grad_(True)
out = tc.cdeint(Xc, func, z0, Xc.interval, adjoint=True, method="dopri5", rtol=1e-3, atol=1e-5, options={"precision": "bf16"})
out.sum().backward()
torch.cuda.synchronize()
print("ok dopri5 adjoint", float(out.abs().max()))
