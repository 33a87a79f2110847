"""Mints tests/golden/gated.pt from the REAL reference (run in the build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_gated_golden.py

MinimalGatedVectorField and GRUGatedVectorField (src/ncde/vector_fields/gating.py:7-61, imported as a package so that the
relative import works) under
torchcde.cdeint, all three vector_field_type modes, adjoint=False: outputs and gradients of loss = sum(out * w).
"""
import os
import sys
import types

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
from src.ncde.vector_fields.gating import GRUGatedVectorField, MinimalGatedVectorField  # noqa: E402

g = torch.Generator().manual_seed(1357)
torch.manual_seed(17)
out = {}
cases = [
    ("min_matmul_lin_rk4", "matmul", 5, 7, 3, 6, 8, 2, "linear", "rk4", 1.0, "grid"),
    ("min_matmul_cub_rk4_half", "matmul", 4, 6, 5, 7, 9, 3, "cubic", "rk4", 0.5, "interval"),
    ("min_matmul_rect_euler", "matmul", 4, 5, 4, 8, 8, 1, "rectilinear", "euler", 1.0, "grid"),
    ("min_eval_lin_rk4", "evaluate", 5, 7, 3, 6, 8, 2, "linear", "rk4", 1.0, "grid"),
    ("min_deriv_cub_rk4", "derivative", 3, 8, 4, 5, 12, 2, "cubic", "rk4", 1.0, "grid"),
    ("min_matmul_wide", "matmul", 6, 5, 33, 32, 16, 2, "linear", "rk4", 1.0, "grid"),
    ("gru_matmul_lin_rk4", "matmul", 5, 7, 3, 6, 8, 2, "linear", "rk4", 1.0, "grid"),
    ("gru_matmul_cub_rk4_half", "matmul", 4, 6, 5, 7, 9, 3, "cubic", "rk4", 0.5, "interval"),
    ("gru_eval_rect_euler", "evaluate", 4, 5, 4, 8, 8, 1, "rectilinear", "euler", 1.0, "grid"),
    ("gru_deriv_lin_rk4", "derivative", 5, 7, 3, 6, 8, 2, "linear", "rk4", 1.0, "grid"),
]
for (name, vft, B, K, C, H, HH, n, interp, method, step, tmode) in cases:
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
    Field = GRUGatedVectorField if name.startswith("gru") else MinimalGatedVectorField
    func = Field(input_dim=C, hidden_dim=H, hidden_hidden_dim=HH, num_layers=n, vector_field_type=vft)
    z0 = (torch.randn(B, H, generator=g) * 0.5).requires_grad_(True)
    t = X.grid_points if tmode == "grid" else X.interval
    w = torch.randn(B, len(t), H, generator=g)
    z = torchcde.cdeint(X, func, z0, t, adjoint=False, vector_field_type=vft, method=method, options={"step_size": step})
    (z * w).sum().backward()
    out[name] = {"coeffs": coeffs, "interp": interp, "vector_field_type": vft, "method": method, "options": {"step_size": step},
                 "t": t, "w": w, "z0": z0.detach().clone(), "dims": {"B": B, "K": K, "C": C, "H": H, "HH": HH, "n": n},
                 "state_dict": {k: v.clone() for k, v in func.state_dict().items()}, "out": z.detach().clone(),
                 "grad_z0": z0.grad.clone(), "grads": {k: p.grad.clone() for k, p in func.named_parameters()},
                 "nfe": func.nfe}
    print(name, tuple(z.shape), func.nfe)
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gated.pt")
torch.save(out, path)
print("wrote", path, os.path.getsize(path))
