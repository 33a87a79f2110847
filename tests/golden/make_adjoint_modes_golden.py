"""Mints tests/golden/adjoint_modes.pt from the REAL reference (run in the build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_adjoint_modes_golden.py

torchcde.cdeint with adjoint=True (torchdiffeq odeint_adjoint, rk4 / euler as forward and adjoint method) for the widened
vector-field modes: vector_field_type evaluate / derivative and the minimal / GRU gated fields (src/ncde/vector_fields).
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
from src.ncde.vector_fields.base import OriginalVectorField  # noqa: E402
from src.ncde.vector_fields.gating import GRUGatedVectorField, MinimalGatedVectorField  # noqa: E402

FIELDS = {"orig": OriginalVectorField, "min": MinimalGatedVectorField, "gru": GRUGatedVectorField}
g = torch.Generator().manual_seed(112233)
torch.manual_seed(29)
out = {}
cases = [
    ("orig_eval_lin_rk4", "orig", "evaluate", 4, 6, 3, 6, 8, 2, "linear", "rk4", 1.0, "grid"),
    ("orig_deriv_cub_rk4_half", "orig", "derivative", 3, 6, 4, 5, 9, 3, "cubic", "rk4", 0.5, "interval"),
    ("min_matmul_lin_rk4", "min", "matmul", 4, 6, 3, 6, 8, 2, "linear", "rk4", 1.0, "grid"),
    ("min_eval_rect_euler", "min", "evaluate", 4, 4, 3, 6, 8, 1, "rectilinear", "euler", 0.5, "grid"),
    ("gru_matmul_lin_rk4", "gru", "matmul", 4, 6, 3, 6, 8, 2, "linear", "rk4", 1.0, "grid"),
    ("gru_deriv_lin_rk4", "gru", "derivative", 4, 6, 3, 6, 8, 2, "linear", "rk4", 1.0, "interval"),
]
for (name, kind, vft, B, K, C, H, HH, n, interp, method, step, tmode) in cases:
    x = torch.randn(B, K, C, generator=g)
    x[..., 0] = torch.arange(K, dtype=torch.float32)
    x[..., 1:] = x[..., 1:].cumsum(-2) * 0.3
    if interp == "linear":
        coeffs = torchcde.linear_interpolation_coeffs(x)
    elif interp == "rectilinear":
        coeffs = torchcde.linear_interpolation_coeffs(x, rectilinear=0)
    else:
        coeffs = torchcde.natural_cubic_coeffs(x)
    X = torchcde.NaturalCubicSpline(coeffs) if interp == "cubic" else torchcde.LinearInterpolation(coeffs)
    func = FIELDS[kind](input_dim=C, hidden_dim=H, hidden_hidden_dim=HH, num_layers=n, vector_field_type=vft)
    z0 = (torch.randn(B, H, generator=g) * 0.5).requires_grad_(True)
    t = X.grid_points if tmode == "grid" else X.interval
    w = torch.randn(B, len(t), H, generator=g)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        z = torchcde.cdeint(X, func, z0, t, adjoint=True, vector_field_type=vft, method=method, options={"step_size": step})
    (z * w).sum().backward()
    out[name] = {"kind": kind, "coeffs": coeffs, "interp": interp, "vector_field_type": vft, "method": method,
                 "options": {"step_size": step}, "t": t, "w": w, "z0": z0.detach().clone(),
                 "dims": {"B": B, "K": K, "C": C, "H": H, "HH": HH, "n": n},
                 "state_dict": {k: v.clone() for k, v in func.state_dict().items()}, "out": z.detach().clone(),
                 "grad_z0": z0.grad.clone(), "grads": {k: p.grad.clone() for k, p in func.named_parameters()}}
    print(name, tuple(z.shape))
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "adjoint_modes.pt")
torch.save(out, path)
print("wrote", path, os.path.getsize(path))
