"""bf16 tensor-core path vs fp32 path vs oracle on a few shapes; prints relative max-norm errors."""
import os, sys, copy
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "tests"), os.path.join(R, "online-neural-cdes_b200")]
import torch
from oracle import cde_oracle as O
import torchcde_b200 as tc

def rel(a, b):
    return float((a.detach().cpu().double() - b.detach().cpu().double()).abs().max() / b.detach().cpu().double().abs().max().clamp_min(1e-30))

def run(B, L, C, H, HH, n, steps_note=""):
    g = torch.Generator().manual_seed(B + L + C)
    x = torch.randn(B, L, C, generator=g); x[..., 0] = torch.arange(L, dtype=torch.float32)
    x[..., 1:] = x[..., 1:].cumsum(-2) * 0.2
    torch.manual_seed(5)
    func = O.SharedMLPField(C, H, HH, n)
    z0 = torch.randn(B, H, generator=g) * 0.5
    cref = O.linear_interpolation_coeffs(x.clone(), rectilinear=0)
    Xr = O.LinearPath(cref)
    w = torch.randn(B, cref.shape[1], H, generator=g)
    z0r = z0.clone().requires_grad_(True)
    oref = O.cdeint(Xr, func, z0r, Xr.grid_points, adjoint=False, method="rk4", options={"step_size": 1})
    (oref * w).sum().backward()
    gref = {k: p.grad.clone() for k, p in func.named_parameters()}
    res = {}
    for prec in ("fp32", "bf16"):
        fd = copy.deepcopy(func).cuda()
        for p in fd.parameters(): p.grad = None
        X = tc.LinearInterpolation(cref.cuda())
        z0d = z0.cuda().requires_grad_(True)
        out = tc.cdeint(X, fd, z0d, X.grid_points, adjoint=False, method="rk4", options={"step_size": 1, "precision": prec})
        (out * w.cuda()).sum().backward()
        torch.cuda.synchronize()
        e = {"out": rel(out, oref), "z0": rel(z0d.grad, z0r.grad)}
        for k, p in fd.named_parameters(): e[k.replace("net_to_hh", "hh").replace("tanh_output_layer", "out")] = rel(p.grad, gref[k])
        res[prec] = e
        print("B=%d L=%d C=%d H=%d HH=%d n=%d %s:" % (B, L, C, H, HH, n, prec), {k: "%.1e" % v for k, v in e.items()}, flush=True)
    return res

if __name__ == "__main__":
    run(1100, 3, 100, 128, 128, 3)
    run(700, 3, 30, 16, 64, 2)
    run(130, 4, 100, 128, 128, 3)
    run(300, 6, 100, 128, 128, 3)
    run(64, 5, 4, 64, 64, 3)
    run(200, 5, 21, 64, 64, 2)
    run(96, 4, 14, 32, 128, 1)
