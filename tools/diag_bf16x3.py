"""Diagnostic: where does the bf16x3 gradient error come from?  cfg-5-shaped problem at full length, ReLU vs tanh hidden layers,
against the oracle in fp32 and fp64; per-row distribution of the z0-gradient error."""
import copy, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "tests"), os.path.join(R, "online-neural-cdes_b200")]
import torch
import bench
import parity_util as PU
import torchcde_b200 as tc
from oracle import cde_oracle as O

B = 128
x, _, _ = bench.synth_batch(B, 7)
cref = O.linear_interpolation_coeffs(x.clone(), rectilinear=0)
g = torch.Generator().manual_seed(4)
z0 = torch.randn(B, 128, generator=g) * 0.5
w = torch.randn(B, 143, 128, generator=g)
for smooth in (False, True):
    torch.manual_seed(3)
    func = O.SharedMLPField(100, 128, 128, 3)
    if smooth:
        mods = [torch.nn.Tanh() if isinstance(m, torch.nn.ReLU) else m for m in func.net_to_hh]
        func.net_to_hh = torch.nn.Sequential(*mods)
    o32 = PU.oracle_solve(func, "linear", cref, z0, w, True)
    o64 = PU.oracle_solve(func, "linear", cref, z0, w, True, dtype=torch.float64)
    for prec in ("fp32", "bf16x3"):
        fd = copy.deepcopy(func).cuda()
        X = tc.LinearInterpolation(cref.cuda())
        z = z0.cuda().requires_grad_(True)
        out = tc.cdeint(X, fd, z, X.grid_points, adjoint=False, method="rk4", options={"step_size": 1, "precision": prec})
        (out * w.cuda()).sum().backward()
        gz = z.grad.cpu().double()
        s64 = o64[1].abs().max()
        row = (gz - o64[1]).abs().amax(1) / s64
        row32 = (o32[1].double() - o64[1]).abs().amax(1) / s64
        pg = {n: PU.rel(p.grad, o64[2][n]) for n, p in fd.named_parameters()}
        print("smooth=%s %s: state vs64 %.1e | z0-grad rows vs fp64: median %.1e p90 %.1e max %.1e (oracle fp32: median %.1e max %.1e) | params %s"
              % (smooth, prec, PU.rel(out, o64[0]), row.median(), row.quantile(0.9), row.max(), row32.median(), row32.max(),
                 {k.split('.')[0][:4] + k[-6:]: "%.0e" % v for k, v in pg.items()}))
