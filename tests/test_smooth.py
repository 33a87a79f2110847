"""SmoothLinearInterpolation (src/ncde/interpolation.py:6-183): linear interpolation with cubic / quintic
gradient-matching regions.  Floating-point work: tolerance 1e-5 relative (fp32), pinned by vectors minted from the real
reference (tests/golden/make_smooth_golden.py) — matching coefficients, evaluate / derivative at scalar times inside and
outside the matching regions, and a cdeint solve over the smoothed path."""
import os

import pytest
import torch

from oracle import cde_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "smooth.pt")


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


class Field(torch.nn.Module):
    def __init__(self, C, H, HH):
        super().__init__()
        self.l1 = torch.nn.Linear(H, HH)
        self.l2 = torch.nn.Linear(HH, H * C)
        self.H, self.C = H, C

    def forward(self, t, z):
        return self.l2(self.l1(z).relu()).tanh().view(-1, self.H, self.C)


def test_oracle_matches_vectors_from_the_real_reference():
    for case in torch.load(GOLDEN):
        X = O.SmoothLinearPath(case["coeffs"], case["eps"], case["quintic"])
        assert rel(X.match, case["match"]) <= 1e-6
        ev = torch.stack([X.evaluate(torch.tensor(t)) for t in case["times"]], 1)
        dv = torch.stack([X.derivative(torch.tensor(t)) for t in case["times"]], 1)
        assert rel(ev, case["evaluate"]) <= 1e-6 and rel(dv, case["derivative"]) <= 1e-6
        func = Field(4, 6, 8)
        func.load_state_dict(case["func"])
        with torch.no_grad():
            sol = O.cdeint(X, func, case["z0"], X.grid_points, adjoint=False, method="rk4", options={"step_size": 0.5})
        assert rel(sol, case["sol"]) <= 1e-5


@pytest.mark.gpu
def test_gpu_matches_vectors_from_the_real_reference():
    from ncde_b200.interpolation import SmoothLinearInterpolation
    import torchcde_b200 as tc
    for case in torch.load(GOLDEN):
        X = SmoothLinearInterpolation(case["coeffs"].cuda(), gradient_matching_eps=case["eps"],
                                      match_second_derivatives=case["quintic"])
        assert rel(X.gradient_matching_coeffs, case["match"]) <= 1e-6
        ev = torch.stack([X.evaluate(torch.tensor(t)) for t in case["times"]], 1)
        dv = torch.stack([X.derivative(torch.tensor(t)) for t in case["times"]], 1)
        assert rel(ev, case["evaluate"]) <= 1e-5 and rel(dv, case["derivative"]) <= 1e-5
        # vectorised evaluation (the reference only takes scalar times) agrees with the scalar calls
        assert torch.equal(X.derivative(torch.tensor(case["times"])), dv)
        func = Field(4, 6, 8)
        func.load_state_dict(case["func"])
        with torch.no_grad():
            sol = tc.cdeint(X, func.cuda(), case["z0"].cuda(), X.grid_points, adjoint=False, method="rk4",
                            options={"step_size": 0.5, "precision": "fp32"})
        assert rel(sol, case["sol"]) <= 1e-5
    with pytest.raises(NotImplementedError):
        tc.cdeint(X, func.cuda(), case["z0"].cuda(), X.grid_points, adjoint=False, method="dopri5")


@pytest.mark.gpu
def test_gpu_smooth_gradients_against_oracle():
    import copy
    from ncde_b200.interpolation import SmoothLinearInterpolation
    import torchcde_b200 as tc
    torch.manual_seed(4)
    B, L, C, H = 9, 8, 5, 16
    x = torch.randn(B, L, C)
    x[..., 0] = torch.arange(L, dtype=torch.float32)
    coeffs = O.linear_interpolation_coeffs(x)
    func = O.SharedMLPField(C, H, H, 2)
    z0 = torch.randn(B, H) * 0.5
    w = torch.randn(B, L, H)
    Xr = O.SmoothLinearPath(coeffs, 0.4, True)
    z0r = z0.clone().requires_grad_(True)
    ref = O.cdeint(Xr, func, z0r, Xr.grid_points, adjoint=False, method="rk4", options={"step_size": 0.25})
    (ref * w).sum().backward()
    fd = copy.deepcopy(func).cuda()
    for p in fd.parameters():
        p.grad = None
    X = SmoothLinearInterpolation(coeffs.cuda(), gradient_matching_eps=0.4, match_second_derivatives=True)
    z0d = z0.cuda().requires_grad_(True)
    out = tc.cdeint(X, fd, z0d, X.grid_points, adjoint=False, method="rk4", options={"step_size": 0.25, "precision": "fp32"})
    (out * w.cuda()).sum().backward()
    assert rel(out, ref) <= 1e-5 and rel(z0d.grad, z0r.grad) <= 1e-5
    for (k, p), (_, q) in zip(fd.named_parameters(), func.named_parameters()):
        assert rel(p.grad, q.grad) <= 1e-5, k
