"""Mints tests/golden/reverse_perturb.pt from the REAL reference (run in the build container only): cdeint with DECREASING output
times (torchdiffeq negates time, _impl/misc.py:262-283) and with options['perturb'] = True (_PerturbFunc nudges the first / last
stage time of every step by one float, misc.py:168-191), fixed-grid solvers, backprop through the solver.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_reverse_golden.py
"""
import importlib.util
import os
import sys
import warnings

import torch

REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(REF, "modules/torchcde"))
sys.path.insert(0, os.path.join(REF, "modules/torchdiffeq"))
import torchcde  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_vf_base", os.path.join(REF, "src/ncde/vector_fields/base.py"))
_vf = importlib.util.module_from_spec(spec)
spec.loader.exec_module(_vf)

g = torch.Generator().manual_seed(1357)
torch.manual_seed(5)
out = {}
cases = [
    # name, interp, method, options, t-mode
    ("rev_lin_rk4_grid", "linear", "rk4", {"step_size": 1}, "rev_grid"),
    ("rev_cub_rk4_half_offgrid", "cubic", "rk4", {"step_size": 0.5}, "rev_offgrid"),
    ("rev_rect_euler_grid", "rectilinear", "euler", {"step_size": 1}, "rev_grid"),
    ("perturb_rect_rk4_grid", "rectilinear", "rk4", {"step_size": 1, "perturb": True}, "grid"),
    ("perturb_lin_euler_grid", "linear", "euler", {"step_size": 1, "perturb": True}, "grid"),
    ("rev_perturb_lin_rk4", "linear", "rk4", {"step_size": 1, "perturb": True}, "rev_grid"),
]
for name, interp, method, options, tmode in cases:
    B, L, C, H, HH, n = 5, 8, 4, 8, 12, 2
    x = torch.randn(B, L, C, generator=g)
    x[..., 0] = torch.arange(L, dtype=torch.float32)
    x[..., 1:] = x[..., 1:].cumsum(-2) * 0.3
    if interp == "rectilinear":
        coeffs = torchcde.linear_interpolation_coeffs(x, rectilinear=0)
    elif interp == "linear":
        coeffs = torchcde.linear_interpolation_coeffs(x)
    else:
        coeffs = torchcde.natural_cubic_coeffs(x)
    X = torchcde.NaturalCubicSpline(coeffs) if interp == "cubic" else torchcde.LinearInterpolation(coeffs)
    func = _vf.OriginalVectorField(input_dim=C, hidden_dim=H, hidden_hidden_dim=HH, num_layers=n)
    z0 = torch.randn(B, H, generator=g) * 0.5
    if tmode == "grid":
        t = X.grid_points
    elif tmode == "rev_grid":
        t = X.grid_points.flip(0)
    else:
        lo, hi = X.interval
        t = torch.cat([lo.view(1), (lo + (hi - lo) * torch.rand(4, generator=g)).sort().values, hi.view(1)]).flip(0)
    w = torch.randn(B, len(t), H, generator=g)
    z0r = z0.clone().requires_grad_(True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        y = torchcde.cdeint(X, func, z0r, t, adjoint=False, method=method, options=dict(options), atol=1e-5, rtol=1e-3)
    (y * w).sum().backward()
    out[name] = {"coeffs": coeffs, "interp": interp, "field": "orig", "dims": {"B": B, "L": L, "C": C, "H": H, "HH": HH, "n": n},
                 "state_dict": {k: v.clone() for k, v in func.state_dict().items()}, "z0": z0, "t": t, "w": w,
                 "kw": dict(adjoint=False, method=method, options=dict(options), atol=1e-5, rtol=1e-3),
                 "out": y.detach().clone(), "grad_z0": z0r.grad.clone(),
                 "grads": {k: p.grad.clone() for k, p in func.named_parameters()}, "nfe": func.nfe}
    print(name, tuple(y.shape), func.nfe)
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reverse_perturb.pt")
torch.save(out, path)
print("wrote", path, os.path.getsize(path))
