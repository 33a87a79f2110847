"""Mints tests/golden/vftype.pt from the REAL reference (run in the build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_vftype_golden.py

torchcde.cdeint with vector_field_type 'evaluate' / 'derivative' (modules/torchcde/torchcde/solver.py:112-137) and the
reference's OriginalVectorField built for those modes (src/ncde/vector_fields/base.py:56-104): outputs and gradients of
loss = sum(out * w), adjoint=False, fixed-grid solvers.
"""
import importlib.util
import os
import sys

import torch

REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(REF, "modules/torchcde"))
sys.path.insert(0, os.path.join(REF, "modules/torchdiffeq"))
import torchcde  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_vf_base", os.path.join(REF, "src/ncde/vector_fields/base.py"))
vf = importlib.util.module_from_spec(spec)
spec.loader.exec_module(vf)

g = torch.Generator().manual_seed(2468)
torch.manual_seed(13)
out = {}
# name, vf type, B, K, C, H, HH, n, interp, method, step, t-mode
cases = [
    ("eval_lin_rk4_grid", "evaluate", 5, 7, 3, 6, 8, 2, "linear", "rk4", 1.0, "grid"),
    ("deriv_lin_rk4_grid", "derivative", 5, 7, 3, 6, 8, 3, "linear", "rk4", 1.0, "grid"),
    ("eval_cub_rk4_half_offgrid", "evaluate", 3, 8, 4, 5, 12, 2, "cubic", "rk4", 0.5, "offgrid"),
    ("deriv_cub_euler_interval", "derivative", 4, 6, 5, 7, 9, 1, "cubic", "euler", 0.25, "interval"),
    ("eval_rect_rk4_grid", "evaluate", 4, 5, 3, 8, 8, 3, "rectilinear", "rk4", 1.0, "grid"),
    ("deriv_lin_rk4_nolayers", "derivative", 4, 6, 3, 5, 8, 0, "linear", "rk4", 1.0, "grid"),
    ("eval_lin_rk4_wide", "evaluate", 6, 5, 40, 32, 16, 2, "linear", "rk4", 1.0, "grid"),
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
    func = vf.OriginalVectorField(input_dim=C, hidden_dim=H, hidden_hidden_dim=HH, num_layers=n, vector_field_type=vft)
    z0 = (torch.randn(B, H, generator=g) * 0.5).requires_grad_(True)
    if tmode == "grid":
        t = X.grid_points
    elif tmode == "interval":
        t = X.interval
    else:
        lo, hi = X.interval
        t = torch.cat([lo.view(1), (lo + (hi - lo) * torch.rand(4, generator=g)).sort().values, hi.view(1)])
    w = torch.randn(B, len(t), H, generator=g)
    z = torchcde.cdeint(X, func, z0, t, adjoint=False, vector_field_type=vft, method=method, options={"step_size": step})
    (z * w).sum().backward()
    out[name] = {"coeffs": coeffs, "interp": interp, "vector_field_type": vft, "method": method, "options": {"step_size": step},
                 "t": t, "w": w, "z0": z0.detach().clone(), "dims": {"B": B, "K": K, "C": C, "H": H, "HH": HH, "n": n},
                 "state_dict": {k: v.clone() for k, v in func.state_dict().items()}, "out": z.detach().clone(),
                 "grad_z0": z0.grad.clone(), "grads": {k: p.grad.clone() for k, p in func.named_parameters()},
                 "nfe": func.nfe}
    print(name, tuple(z.shape), func.nfe)
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vftype.pt")
torch.save(out, path)
print("wrote", path, os.path.getsize(path))
