"""All-tensor-core path (hidden layers on tcgen05 too) vs the bf16 path with CUDA-core hidden layers vs fp32: forward only."""
import os, sys, copy
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "tests"), os.path.join(R, "online-neural-cdes_b200")]
import torch
from oracle import cde_oracle as O
import torchcde_b200 as tc

def rel(a, b):
    return float((a.detach().cpu().double() - b.detach().cpu().double()).abs().max() / b.detach().cpu().double().abs().max().clamp_min(1e-30))

def run(B, L, C, H, HH, n, method="rk4"):
    g = torch.Generator().manual_seed(B + L + C)
    x = torch.randn(B, L, C, generator=g); x[..., 0] = torch.arange(L, dtype=torch.float32)
    x[..., 1:] = x[..., 1:].cumsum(-2) * 0.2
    torch.manual_seed(5)
    func = O.SharedMLPField(C, H, HH, n) if n > 0 else O.ToyField(C, H, width=HH)
    z0 = torch.randn(B, H, generator=g) * 0.5
    cref = O.linear_interpolation_coeffs(x.clone(), rectilinear=0)
    fd = copy.deepcopy(func).cuda()
    X = tc.LinearInterpolation(cref.cuda())
    with torch.no_grad():
        outs = {}
        for prec in ("fp32", "bf16"):
            outs[prec] = tc.cdeint(X, fd, z0.cuda(), X.grid_points, adjoint=False, method=method, options={"step_size": 1, "precision": prec})
        torch.cuda.synchronize()
    print("B=%d L=%d C=%d H=%d HH=%d n=%d %s: bf16 vs fp32 %.2e  finite %s" % (B, L, C, H, HH, n, method, rel(outs["bf16"], outs["fp32"]),
          bool(torch.isfinite(outs["bf16"]).all())), flush=True)

if __name__ == "__main__":
    for args in [(1100, 3, 100, 128, 128, 3), (700, 4, 30, 16, 64, 2), (130, 4, 100, 128, 128, 3), (64, 5, 4, 64, 64, 3), (96, 4, 14, 32, 128, 1), (33, 3, 2, 32, 128, 0)]:
        run(*args)
    run(300, 5, 21, 64, 64, 2, method="euler")
