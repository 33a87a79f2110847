"""Mints tests/golden/pathgrad.pt from the REAL reference (run in the build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_pathgrad_golden.py

Gradients w.r.t. the control-path coefficients through torchcde.cdeint (adjoint=False) — what stacked Neural CDEs rely on
(modules/torchcde/test/test_tricks.py:54-106) — the backward of evaluate / derivative, and an end-to-end
StackedNeuralCDE (src/ncde/stacked.py, imported with `autots` stubbed: only attention.py needs it).
"""
import os
import sys
import types
import warnings

import torch

REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(REF, "modules/torchcde"))
sys.path.insert(0, os.path.join(REF, "modules/torchdiffeq"))
sys.path.insert(0, REF)
_a = types.ModuleType("autots")
_p = types.ModuleType("autots.preprocessing")
_p.ForwardFill = _p.PadRaggedTensors = _p.SimplePipeline = object
_a.preprocessing = _p
sys.modules["autots"] = _a
sys.modules["autots.preprocessing"] = _p
import torchcde  # noqa: E402
from src.ncde import StackedNeuralCDE  # noqa: E402
from src.ncde.vector_fields.base import OriginalVectorField  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
g = torch.Generator().manual_seed(777)
torch.manual_seed(5)
out = {"cdeint": {}, "eval": {}, "stacked": {}}

# name, B, K, C, H, HH, n, interp, method, step, t-mode, uniform knots
cases = [
    ("lin_rk4_grid", 5, 7, 3, 6, 8, 2, "linear", "rk4", 1.0, "grid", True),
    ("lin_rk4_half_interval", 4, 6, 5, 7, 9, 3, "linear", "rk4", 0.5, "interval", True),
    ("lin_euler_nonuniform", 3, 8, 4, 8, 8, 1, "linear", "euler", None, "grid", False),
    ("cub_rk4_offgrid", 3, 8, 3, 5, 12, 2, "cubic", "rk4", 0.5, "offgrid", True),
    ("lin_rk4_wide", 6, 5, 33, 32, 16, 2, "linear", "rk4", 1.0, "grid", True),
]
for (name, B, K, C, H, HH, n, interp, method, step, tmode, uniform) in cases:
    x = torch.randn(B, K, C, generator=g)
    x[..., 1:] = x[..., 1:].cumsum(-2) * 0.3
    knots = torch.arange(K, dtype=torch.float32) if uniform else torch.rand(K, generator=g).cumsum(0) * 2
    x[..., 0] = knots
    if interp == "linear":
        coeffs = torchcde.linear_interpolation_coeffs(x, knots).clone()
    else:
        coeffs = torchcde.natural_cubic_coeffs(x, knots).clone()
    coeffs.requires_grad_(True)
    X = torchcde.LinearInterpolation(coeffs, knots) if interp == "linear" else torchcde.NaturalCubicSpline(coeffs, knots)
    func = OriginalVectorField(input_dim=C, hidden_dim=H, hidden_hidden_dim=HH, num_layers=n)
    z0 = (torch.randn(B, H, generator=g) * 0.5).requires_grad_(True)
    if tmode == "grid":
        t = X.grid_points
    elif tmode == "interval":
        t = X.interval
    else:
        lo, hi = X.interval
        t = torch.cat([lo.view(1), (lo + (hi - lo) * torch.rand(4, generator=g)).sort().values, hi.view(1)])
    w = torch.randn(B, len(t), H, generator=g)
    options = {} if step is None else {"step_size": step}
    z = torchcde.cdeint(X, func, z0, t, adjoint=False, method=method, options=options)
    (z * w).sum().backward()
    out["cdeint"][name] = {
        "coeffs": coeffs.detach().clone(), "knots": knots, "interp": interp, "method": method, "options": options, "t": t.detach(),
        "w": w, "z0": z0.detach().clone(), "dims": {"B": B, "K": K, "C": C, "H": H, "HH": HH, "n": n},
        "state_dict": {k: v.clone() for k, v in func.state_dict().items()}, "out": z.detach().clone(),
        "grad_coeffs": coeffs.grad.clone(), "grad_z0": z0.grad.clone(),
        "grads": {k: p.grad.clone() for k, p in func.named_parameters()},
    }
    print(name, tuple(z.shape), float(coeffs.grad.abs().max()))

# backward of evaluate / derivative
for interp in ("linear", "cubic"):
    for dtype in (torch.float32, torch.float64):
        B, K, C = 4, 9, 3
        knots = (torch.rand(K, generator=g, dtype=torch.float64).cumsum(0) * 1.5).to(dtype)
        x = torch.randn(B, K, C, generator=g, dtype=torch.float64).to(dtype)
        coeffs = (torchcde.linear_interpolation_coeffs(x, knots) if interp == "linear"
                  else torchcde.natural_cubic_coeffs(x, knots)).clone().requires_grad_(True)
        X = torchcde.LinearInterpolation(coeffs, knots) if interp == "linear" else torchcde.NaturalCubicSpline(coeffs, knots)
        tq = torch.cat([knots[:1], knots[3:4], knots[-1:], knots[0] + (knots[-1] - knots[0]) * torch.rand(6, generator=g, dtype=torch.float64).to(dtype)])
        rec = {"coeffs": coeffs.detach().clone(), "knots": knots, "tq": tq, "interp": interp}
        for which in ("evaluate", "derivative"):
            coeffs.grad = None
            val = getattr(X, which)(tq)
            w = torch.randn(val.shape, generator=g, dtype=torch.float64).to(dtype)
            (val * w).sum().backward()
            rec[which] = {"w": w, "val": val.detach().clone(), "grad": coeffs.grad.clone()}
        out["eval"]["{}_{}".format(interp, str(dtype).split(".")[-1])] = rec

# stacked Neural CDE end to end (adjoint=False so that gradients flow through the intermediate paths)
for name, static_dim in (("plain", None), ("static_all", 3)):
    B, L, C = 4, 6, 3
    x = torch.randn(B, L, C, generator=g)
    x[..., 0] = torch.arange(L, dtype=torch.float32)
    x[..., 1:] = x[..., 1:].cumsum(-2) * 0.3
    coeffs = torchcde.linear_interpolation_coeffs(x).clone().requires_grad_(True)
    model = StackedNeuralCDE(C, [5, 6], 2, hidden_hidden_dim=7, static_dim=static_dim, adjoint=False, return_sequences=True,
                             static_in_all_layers=static_dim is not None)
    static = torch.randn(B, static_dim, generator=g) if static_dim else None
    inputs = coeffs if static is None else [static, coeffs]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        y = model(inputs)
    w = torch.randn(y.shape, generator=g)
    (y * w).sum().backward()
    out["stacked"][name] = {
        "coeffs": coeffs.detach().clone(), "static": static, "w": w, "out": y.detach().clone(), "grad_coeffs": coeffs.grad.clone(),
        "state_dict": {k: v.clone() for k, v in model.state_dict().items()},
        "grads": {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None},
        "args": {"input_dim": C, "hidden_dims": [5, 6], "output_dim": 2, "hidden_hidden_dim": 7, "static_dim": static_dim},
    }
    print("stacked", name, tuple(y.shape), sorted(out["stacked"][name]["grads"])[:4])

path = os.path.join(HERE, "pathgrad.pt")
torch.save(out, path)
print("wrote", path, os.path.getsize(path))
