"""Side-by-side step traces (dt, error ratio, accepted) of the oracle and the CUDA path for a dopri5 + adjoint solve."""
import os, sys, copy
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "online-neural-cdes_b200")]
import torch
from oracle import cde_oracle as O
import torchcde_b200 as tc
from torchcde_b200 import adaptive


def case(B, L, C, H, HH, n, t_mode, kw, path="cubic", nshow=12):
    g = torch.Generator().manual_seed(B + L); torch.manual_seed(6)
    x = torch.randn(B, L, C, generator=g); x[..., 0] = torch.arange(L, dtype=torch.float32); x[..., 1:] = x[..., 1:].cumsum(-2) * 0.15
    func = O.SharedMLPField(C, H, HH, n); z0 = torch.randn(B, H, generator=g) * 0.5
    if path == "cubic":
        cref = O.natural_cubic_coeffs(x); Xr = O.CubicPath(cref); X = tc.NaturalCubicSpline(cref.cuda())
    else:
        cref = O.linear_interpolation_coeffs(x); Xr = O.LinearPath(cref); X = tc.LinearInterpolation(cref.cuda())
    t = Xr.interval if t_mode == "interval" else Xr.grid_points[::3].contiguous()
    w = torch.randn(B, len(t), H, generator=g)
    fr = copy.deepcopy(func); z0r = z0.clone().requires_grad_(True); st = {'record_norms': True}
    o = O.cdeint(Xr, fr, z0r, t, adjoint=True, method="dopri5", stats=st, **kw)
    nf = st["attempted"]
    (o * w).sum().backward()
    fd = copy.deepcopy(func).cuda(); z0d = z0.cuda().requires_grad_(True)
    og = tc.cdeint(X, fd, z0d, t.cuda(), adjoint=True, method="dopri5", **kw)
    (og * w.cuda()).sum().backward()
    fs, bs = dict(adaptive.last_stats), dict(adaptive.last_adjoint_stats)
    print(path, t_mode, kw)
    print("  forward : oracle attempted %d | gpu attempted %d" % (nf, fs["attempted"]))
    for i, (a, b) in enumerate(zip(st["trace"][:nf], fs["trace"])):
        flag = "" if (a[2] == bool(int(b[2]) % 10)) and abs(a[0] - b[0]) <= 1e-3 * abs(a[0]) else "   <<<"
        if i < nshow: print("    %3d  oracle dt %.6f ratio %10.4e acc %d | gpu dt %.6f ratio %10.4e acc %d%s" % (i, a[0], a[1], a[2], b[0], b[1], b[2], flag))
    print("  backward: oracle attempted %d | gpu attempted %d" % (st["attempted"] - nf, bs["attempted"]))
    for i, (a, b) in enumerate(zip(st["trace"][nf:], bs["trace"])):
        flag = "" if (a[2] == bool(int(b[2]) % 10)) and abs(a[0] - b[0]) <= 1e-3 * abs(a[0]) else "   <<<"
        if i < nshow: print("    %3d  oracle dt %.6f ratio %10.4e acc %d | gpu dt %.6f ratio %10.4e acc %d seg %d%s   oracle norms %s" % (i, a[0], a[1], a[2], b[0], b[1], int(b[2]) % 10, int(b[2]) // 10, flag, ["%.2e" % v for v in st["norm_vals"][i]]))


if __name__ == "__main__":
    case(16, 40, 5, 16, 16, 3, "interval", dict(rtol=1e-3, atol=1e-5, options={"min_step": 0.5, "first_step": 0.5}))
    case(6, 10, 4, 8, 8, 2, "online", dict(rtol=1e-4, atol=1e-6, options={"first_step": 0.05}), nshow=12)
