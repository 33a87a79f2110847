import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "online-neural-cdes_b200")]
import torch
from oracle import cde_oracle as O
import torchcde_b200 as tc
from torchcde_b200 import adaptive
torch.manual_seed(0)
x = torch.rand(3, 8, 2)
func = O.SharedMLPField(2, 4, 8, 1)
z0 = torch.rand(3, 4)
Xr = O.CubicPath(O.natural_cubic_coeffs(x))
X = tc.NaturalCubicSpline(tc.natural_cubic_coeffs(x.cuda()))
fd = O.SharedMLPField(2, 4, 8, 1).cuda(); fd.load_state_dict(func.state_dict())
def both(t, **kw):
    t = torch.tensor(t, dtype=torch.float64)
    with torch.no_grad():
        st = {}
        ref = O.cdeint(Xr, func, z0, t, adjoint=False, method="dopri5", stats=st, **kw)
        out = tc.cdeint(X, fd, z0.cuda(), t.cuda(), adjoint=False, method="dopri5", **kw).cpu()
    err = (out - ref).abs().amax(dim=(0, 2)) / ref.abs().max()
    print(kw, "ref stats", st, "gpu", {k: adaptive.last_stats[k] for k in ("attempted", "accepted", "nfe", "first_step", "init_h0_d0_d1_d2")})
    print("   per-output rel err:", ["%.1e" % e for e in err.tolist()])
import oracle.cde_oracle as OO
_orig = OO._initial_step
def spy(f, t0, y0, order, rtol, atol, norm, f0):
    r = _orig(f, t0, y0, order, rtol, atol, norm, f0)
    scale = atol + torch.abs(y0) * rtol
    print("   ORACLE first step", float(r), "d0", float(norm(y0 / scale)), "d1", float(norm(f0 / scale)))
    return r
OO._initial_step = spy
_o2 = OO._next_step_size
def spy2(dt, ratio, *a):
    print("   ORACLE dt %.9g ratio %.9g" % (float(dt), float(ratio)))
    return _o2(dt, ratio, *a)
OO._next_step_size = spy2
both([0., 2.0], rtol=1e-4, atol=1e-6, options={})
for r in adaptive.last_stats["trace"]: print("   GPU    dt %.9g ratio %.9g acc %d" % tuple(r))
