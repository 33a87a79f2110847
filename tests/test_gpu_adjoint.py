"""adjoint=True with fixed-grid methods: the continuous adjoint integrated by ncde_solve_adjoint_bwd against the golden
vectors of the real reference (lin_rk4_adjoint) and against the oracle's restatement of adjoint.py."""
import copy

import pytest
import torch

from oracle import cde_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_golden_rk4_adjoint(golden_cdeint):
    import torchcde_b200 as tc
    rec = golden_cdeint["lin_rk4_adjoint"]
    d = rec["dims"]
    func = O.SharedMLPField(d["C"], d["H"], d["HH"], d["n"])
    func.load_state_dict(rec["state_dict"])
    func = func.cuda()
    X = tc.LinearInterpolation(rec["coeffs"].cuda())
    z0 = rec["z0"].cuda().requires_grad_(True)
    kw = rec["kw"]
    assert kw["adjoint"] is True
    out = tc.cdeint(X, func, z0, rec["t"].cuda(), adjoint=True, method="rk4", rtol=kw["rtol"], atol=kw["atol"],
                    options=dict(kw["options"]))
    (out * rec["w"].cuda()).sum().backward()
    assert rel(out, rec["out"]) <= TOL
    assert rel(z0.grad, rec["grad_z0"]) <= TOL
    for n, p in func.named_parameters():
        assert rel(p.grad, rec["grads"][n]) <= TOL, n


@pytest.mark.parametrize("case", ["cubic_half_step_offgrid", "rect_euler", "cfg2_shape"])
def test_adjoint_against_oracle(case):
    import torchcde_b200 as tc
    g = torch.Generator().manual_seed(3)
    torch.manual_seed(4)
    if case == "cubic_half_step_offgrid":
        B, L, C, H, HH, n, method, step = 5, 8, 3, 8, 12, 2, "rk4", 0.5
    elif case == "rect_euler":
        B, L, C, H, HH, n, method, step = 9, 6, 4, 16, 16, 3, "euler", 1
    else:
        B, L, C, H, HH, n, method, step = 130, 20, 4, 64, 64, 3, "rk4", 1
    x = torch.randn(B, L, C, generator=g)
    x[..., 0] = torch.arange(L, dtype=torch.float32)
    x[..., 1:] = x[..., 1:].cumsum(-2) * 0.2
    func = O.SharedMLPField(C, H, HH, n)
    z0 = torch.randn(B, H, generator=g) * 0.5
    if case == "cubic_half_step_offgrid":
        cref = O.natural_cubic_coeffs(x)
        Xr = O.CubicPath(cref)
        lo, hi = Xr.interval
        t = torch.cat([lo.view(1), (lo + (hi - lo) * torch.rand(4, generator=g)).sort().values, hi.view(1)])
    elif case == "rect_euler":
        cref = O.linear_interpolation_coeffs(x.clone(), rectilinear=0)
        Xr = O.LinearPath(cref)
        t = Xr.grid_points
    else:
        cref = O.linear_interpolation_coeffs(x.clone())
        Xr = O.LinearPath(cref)
        t = Xr.interval
    w = torch.randn(B, len(t), H, generator=g)
    z0r = z0.clone().requires_grad_(True)
    oref = O.cdeint(Xr, func, z0r, t, adjoint=True, method=method, options={"step_size": step})
    (oref * w).sum().backward()
    gref = {k: p.grad.clone() for k, p in func.named_parameters()}
    fd = copy.deepcopy(func).cuda()
    for p in fd.parameters():
        p.grad = None
    X = tc.NaturalCubicSpline(cref.cuda()) if case == "cubic_half_step_offgrid" else tc.LinearInterpolation(cref.cuda())
    z0d = z0.cuda().requires_grad_(True)
    out = tc.cdeint(X, fd, z0d, t.cuda(), adjoint=True, method=method, options={"step_size": step})
    (out * w.cuda()).sum().backward()
    errs = {"out": rel(out, oref), "z0": rel(z0d.grad, z0r.grad)}
    for k, p in fd.named_parameters():
        errs[k] = rel(p.grad, gref[k])
    assert max(errs.values()) <= TOL, errs


def test_adjoint_differs_from_backprop_by_truncation_error_only():
    """Sanity: continuous adjoint and discretise-then-optimise gradients agree to O(h^4) but are not identical."""
    import torchcde_b200 as tc
    torch.manual_seed(0)
    x = torch.rand(4, 9, 3).cuda()
    X = tc.NaturalCubicSpline(tc.natural_cubic_coeffs(x))
    func = O.SharedMLPField(3, 8, 8, 2).cuda()
    grads = {}
    for adjoint in (True, False):
        func.zero_grad()
        z0 = (torch.ones(4, 8) * 0.3).cuda().requires_grad_(True)
        out = tc.cdeint(X, func, z0, X.interval, adjoint=adjoint, method="rk4", options={"step_size": 0.25})
        out[:, -1].sum().backward()
        grads[adjoint] = z0.grad.clone()
    r = rel(grads[True], grads[False])
    assert 0 < r < 5e-2
