"""Mints tests/golden/neuralcde.pt from the REAL reference model wrapper src/ncde/ncde.py (run in the build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_neuralcde_golden.py

`src.ncde` is imported as a package with `autots` stubbed (only attention.py needs it).  Each case: NeuralCDE(...) forward on
seeded coefficients (+ static features), loss = sum(out * w), gradients of every parameter.
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
from src.ncde import NeuralCDE  # noqa: E402

g = torch.Generator().manual_seed(8642)
torch.manual_seed(23)
out = {}
# name: (constructor kwargs, interpolation of the data, B, L, C)
cases = {
    "orig_rect_static_online": (dict(static_dim=4, hidden_hidden_dim=9, num_layers=3, interpolation="rectilinear", adjoint=False,
                                     return_sequences=True), 5, 6, 4),
    "orig_cubic_adjoint_terminal": (dict(hidden_hidden_dim=8, num_layers=2, interpolation="cubic", adjoint=True), 4, 7, 3),
    "orig_linear_noinitial": (dict(use_initial=False, hidden_hidden_dim=8, num_layers=2, interpolation="linear", adjoint=False,
                                   return_sequences=True, apply_final_linear=False), 4, 6, 3),
    "minimal_evaluate_online": (dict(hidden_hidden_dim=8, num_layers=2, interpolation="linear", adjoint=False,
                                     vector_field="minimal", vector_field_type="evaluate", return_sequences=True), 5, 6, 3),
    "gru_derivative_static_terminal": (dict(static_dim=3, hidden_hidden_dim=8, num_layers=3, interpolation="linear",
                                            adjoint=False, vector_field="gru", vector_field_type="derivative"), 4, 6, 4),
    "gru_matmul_rect_online": (dict(hidden_hidden_dim=10, num_layers=2, interpolation="rectilinear", adjoint=False,
                                    vector_field="gru", return_sequences=True), 4, 5, 3),
    "orig_smooth_cubic": (dict(hidden_hidden_dim=8, num_layers=2, interpolation="linear_cubic_smoothing",
                               interpolation_eps=0.4, adjoint=False, return_sequences=True), 4, 6, 3),
}
for name, (kw, B, L, C) in cases.items():
    H, O_ = 6, 2
    x = torch.randn(B, L, C, generator=g)
    x[..., 0] = torch.arange(L, dtype=torch.float32)
    x[..., 1:] = x[..., 1:].cumsum(-2) * 0.3
    interp = kw["interpolation"]
    if interp == "rectilinear":
        coeffs = torchcde.linear_interpolation_coeffs(x, rectilinear=0)
    elif interp == "cubic":
        coeffs = torchcde.natural_cubic_coeffs(x)
    else:
        coeffs = torchcde.linear_interpolation_coeffs(x)
    model = NeuralCDE(C, H, O_, **kw)
    static = torch.randn(B, kw["static_dim"], generator=g) if kw.get("static_dim") else None
    inputs = coeffs if static is None else (static, coeffs)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        y = model(inputs)
    w = torch.randn(y.shape, generator=g)
    (y * w).sum().backward()
    out[name] = {"kwargs": kw, "dims": (C, H, O_), "coeffs": coeffs, "static": static, "w": w, "out": y.detach().clone(),
                 "state_dict": {k: v.clone() for k, v in model.state_dict().items()},
                 "grads": {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}, "nfe": model.nfe}
    print(name, tuple(y.shape), model.nfe, len(out[name]["grads"]))
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "neuralcde.pt")
torch.save(out, path)
print("wrote", path, os.path.getsize(path))
