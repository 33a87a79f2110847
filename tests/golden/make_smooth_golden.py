"""Mints tests/golden/smooth.pt from the REAL reference (src/ncde/interpolation.py loaded by path; run in the build container:
    PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference/modules/torchcde:/root/reference/modules/torchdiffeq python tests/golden/make_smooth_golden.py
SmoothLinearInterpolation evaluate / derivative at scalar times, and a cdeint solve (rk4, step 1/2) over the smoothed path."""
import importlib.util
import os
import torch
import torchcde

spec = importlib.util.spec_from_file_location("ref_interpolation", "/root/reference/src/ncde/interpolation.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


class Field(torch.nn.Module):
    def __init__(self, C, H, HH):
        super().__init__()
        self.l1 = torch.nn.Linear(H, HH)
        self.l2 = torch.nn.Linear(HH, H * C)
        self.H, self.C = H, C

    def forward(self, t, z):
        return self.l2(self.l1(z).relu()).tanh().view(-1, self.H, self.C)


torch.manual_seed(0)
cases = []
for quintic, eps in [(False, 0.3), (True, 0.5), (False, 1.0)]:
    B, L, C, H = 5, 7, 4, 6
    x = torch.randn(B, L, C)
    x[..., 0] = torch.arange(L, dtype=torch.float32)
    coeffs = torchcde.linear_interpolation_coeffs(x)
    X = ref.SmoothLinearInterpolation(coeffs, gradient_matching_eps=eps, match_second_derivatives=quintic)
    times = [0.0, 0.2, 1.0, 1.1, 1.25, 2.0, 2.4, 3.05, 4.5, 5.0, 5.29, 6.0]
    ev = torch.stack([X.evaluate(torch.tensor(t)) for t in times], 1)
    dv = torch.stack([X.derivative(torch.tensor(t)) for t in times], 1)
    func = Field(C, H, 8)
    z0 = torch.randn(B, H) * 0.5
    with torch.no_grad():
        sol = torchcde.cdeint(X, func, z0, X.grid_points, adjoint=False, method="rk4", options={"step_size": 0.5})
    cases.append({"coeffs": coeffs, "eps": eps, "quintic": quintic, "times": times, "evaluate": ev, "derivative": dv,
                  "match": X.gradient_matching_coeffs, "func": func.state_dict(), "z0": z0, "sol": sol})
torch.save(cases, os.path.join(os.path.dirname(os.path.abspath(__file__)), "smooth.pt"))
print([tuple(c["sol"].shape) for c in cases])
