"""bf16 vs fp32 path at a large per-GPU batch (8192 series, cfg-5 widths), forward + backward: sizes / index ranges sanity."""
import os, sys, copy
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "online-neural-cdes_b200")]
import torch
from oracle import cde_oracle as O
import torchcde_b200 as tc

def rel(a, b):
    return float((a.detach().double() - b.detach().double()).abs().max() / b.detach().double().abs().max().clamp_min(1e-30))

torch.manual_seed(0)
B, L, C, H = 8192, 5, 100, 128
x = torch.randn(B, L, C); x[..., 0] = torch.arange(L, dtype=torch.float32); x[..., 1:] = x[..., 1:].cumsum(-2) * 0.2
coeffs = tc.linear_interpolation_coeffs(x.cuda(), rectilinear=0)
X = tc.LinearInterpolation(coeffs)
func = O.SharedMLPField(C, H, H, 3)
z0 = torch.randn(B, H) * 0.5
w = torch.randn(B, coeffs.shape[1], H).cuda()
res = {}
for prec in ("fp32", "bf16"):
    fd = copy.deepcopy(func).cuda()
    z = z0.cuda().requires_grad_(True)
    out = tc.cdeint(X, fd, z, X.grid_points, adjoint=False, method="rk4", options={"step_size": 1, "precision": prec})
    (out * w).sum().backward()
    torch.cuda.synchronize()
    res[prec] = (out.detach(), z.grad, {k: p.grad for k, p in fd.named_parameters()})
o32, g32, p32 = res["fp32"]; o16, g16, p16 = res["bf16"]
print("out %.2e z0 %.2e" % (rel(o16, o32), rel(g16, g32)), {k[-12:]: "%.1e" % rel(p16[k], p32[k]) for k in p32})
print("peak memory GB", torch.cuda.max_memory_allocated() / 2**30)
