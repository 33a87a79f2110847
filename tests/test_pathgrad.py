"""Gradients with respect to the control path (stacked Neural CDEs).

Golden vectors: tests/golden/pathgrad.pt, minted from the REAL reference by tests/golden/make_pathgrad_golden.py
(torchcde.cdeint with adjoint=False and coefficients that require gradients; evaluate / derivative backward;
src/ncde/stacked.py end to end).  CPU tests pin the oracle against them; GPU tests compare the CUDA path (through the C ABI:
ncde_solve_bwd's grad_coeffs, ncde_path_eval_bwd) with the golden vectors and with the oracle on config-shaped problems.

Tolerance: relative max-norm 1e-5 for fp32 (BASELINE.json north_star), 1e-12 for the fp64 evaluate / derivative backward.
"""
import os
import warnings

import pytest
import torch

from oracle import cde_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pathgrad.pt")
TOL = 1e-5


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLDEN)


def rel(a, b):
    a = a.detach().cpu()
    b = b.detach().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _field(rec):
    d = rec["dims"]
    f = O.SharedMLPField(d["C"], d["H"], d["HH"], d["n"])
    f.load_state_dict(rec["state_dict"])
    return f


CDEINT_CASES = ["lin_rk4_grid", "lin_rk4_half_interval", "lin_euler_nonuniform", "cub_rk4_offgrid", "lin_rk4_wide"]


# ---------------------------------------------------------------------------------------------------------------
# CPU: the oracle reproduces the reference's path gradients (pins the oracle for this row)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", CDEINT_CASES)
def test_oracle_path_gradient_matches_reference(gold, name):
    rec = gold["cdeint"][name]
    func = _field(rec)
    coeffs = rec["coeffs"].clone().requires_grad_(True)
    X = O.CubicPath(coeffs, rec["knots"]) if rec["interp"] == "cubic" else O.LinearPath(coeffs, rec["knots"])
    z0 = rec["z0"].clone().requires_grad_(True)
    out = O.cdeint(X, func, z0, rec["t"], adjoint=False, method=rec["method"], options=dict(rec["options"]))
    (out * rec["w"]).sum().backward()
    assert rel(out, rec["out"]) <= 1e-6
    assert rel(coeffs.grad, rec["grad_coeffs"]) <= 1e-5
    assert rel(z0.grad, rec["grad_z0"]) <= 1e-5


@pytest.mark.parametrize("key", ["linear_float32", "linear_float64", "cubic_float32", "cubic_float64"])
def test_oracle_eval_backward_matches_reference(gold, key):
    rec = gold["eval"][key]
    for which in ("evaluate", "derivative"):
        coeffs = rec["coeffs"].clone().requires_grad_(True)
        X = O.CubicPath(coeffs, rec["knots"]) if rec["interp"] == "cubic" else O.LinearPath(coeffs, rec["knots"])
        val = getattr(X, which)(rec["tq"])
        (val * rec[which]["w"]).sum().backward()
        tol = 1e-6 if coeffs.dtype == torch.float32 else 1e-13
        assert rel(val, rec[which]["val"]) <= tol
        assert rel(coeffs.grad, rec[which]["grad"]) <= tol


def test_stacked_state_dict_is_interchangeable_with_the_reference(gold):
    import ncde_b200
    for name, rec in gold["stacked"].items():
        a = rec["args"]
        m = ncde_b200.StackedNeuralCDE(a["input_dim"], a["hidden_dims"], a["output_dim"], hidden_hidden_dim=a["hidden_hidden_dim"],
                                       static_dim=a["static_dim"], adjoint=False, return_sequences=True,
                                       static_in_all_layers=a["static_dim"] is not None)
        missing, unexpected = m.load_state_dict(rec["state_dict"])
        assert not missing and not unexpected


# ---------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def tc():
    import torchcde_b200
    assert torch.cuda.is_available()
    return torchcde_b200


def _run_cuda(tc, rec, **extra):
    func = _field(rec).cuda()
    coeffs = rec["coeffs"].cuda().requires_grad_(True)
    knots = rec["knots"].cuda()
    X = tc.NaturalCubicSpline(coeffs, knots) if rec["interp"] == "cubic" else tc.LinearInterpolation(coeffs, knots)
    z0 = rec["z0"].cuda().requires_grad_(True)
    opts = dict(rec["options"])
    opts.update(extra)
    out = tc.cdeint(X, func, z0, rec["t"].cuda(), adjoint=False, method=rec["method"], options=opts)
    (out * rec["w"].cuda()).sum().backward()
    torch.cuda.synchronize()
    return out, coeffs.grad, z0.grad, {n: p.grad for n, p in func.named_parameters()}


@pytest.mark.gpu
@pytest.mark.parametrize("name", CDEINT_CASES)
def test_golden_path_gradient(tc, gold, name):
    rec = gold["cdeint"][name]
    out, gc, gz0, grads = _run_cuda(tc, rec)
    assert rel(out, rec["out"]) <= TOL
    assert gc.shape == rec["grad_coeffs"].shape
    assert rel(gc, rec["grad_coeffs"]) <= TOL
    assert rel(gz0, rec["grad_z0"]) <= TOL
    for n, g in rec["grads"].items():
        assert rel(grads[n], g) <= TOL, n


@pytest.mark.gpu
@pytest.mark.parametrize("key", ["linear_float32", "linear_float64", "cubic_float32", "cubic_float64"])
def test_golden_eval_backward(tc, gold, key):
    rec = gold["eval"][key]
    for which in ("evaluate", "derivative"):
        coeffs = rec["coeffs"].cuda().requires_grad_(True)
        knots = rec["knots"].cuda()
        X = tc.NaturalCubicSpline(coeffs, knots) if rec["interp"] == "cubic" else tc.LinearInterpolation(coeffs, knots)
        val = getattr(X, which)(rec["tq"].cuda())
        assert val.requires_grad
        (val * rec[which]["w"].cuda()).sum().backward()
        tol = 1e-6 if coeffs.dtype == torch.float32 else 1e-12
        assert rel(val, rec[which]["val"]) <= tol
        assert rel(coeffs.grad, rec[which]["grad"]) <= tol


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["plain", "static_all"])
def test_golden_stacked_neural_cde(gold, name):
    """src/ncde/stacked.py end to end: outputs and every gradient — incl. the input path's and the first link's parameters',
    which only exist if the gradient crosses the intermediate path — against the real reference."""
    import ncde_b200
    rec = gold["stacked"][name]
    a = rec["args"]
    m = ncde_b200.StackedNeuralCDE(a["input_dim"], a["hidden_dims"], a["output_dim"], hidden_hidden_dim=a["hidden_hidden_dim"],
                                   static_dim=a["static_dim"], adjoint=False, return_sequences=True,
                                   static_in_all_layers=a["static_dim"] is not None).cuda()
    m.load_state_dict(rec["state_dict"])
    coeffs = rec["coeffs"].cuda().requires_grad_(True)
    inputs = coeffs if rec["static"] is None else [rec["static"].cuda(), coeffs]
    y = m(inputs)
    (y * rec["w"].cuda()).sum().backward()
    torch.cuda.synchronize()
    assert rel(y, rec["out"]) <= TOL
    assert rel(coeffs.grad, rec["grad_coeffs"]) <= TOL
    got = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    assert sorted(got) == sorted(rec["grads"])
    for k, g in rec["grads"].items():
        assert rel(got[k], g) <= TOL, k


def _oracle_pair(tc, B, K, C, H, HH, n, interp, seed, rows=None):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, K, C, generator=g).cumsum(-2) * 0.2
    torch.manual_seed(seed)
    func = O.SharedMLPField(C, H, HH, n)
    z0 = torch.randn(B, H, generator=g) * 0.5
    cref = (O.natural_cubic_coeffs(x) if interp == "cubic" else x.clone())
    w = torch.randn(B, K, H, generator=g)
    res = {}
    if rows is None:
        c = cref.clone().requires_grad_(True)
        Xr = O.CubicPath(c) if interp == "cubic" else O.LinearPath(c)
        out = O.cdeint(Xr, func, z0, Xr.grid_points, adjoint=False, method="rk4", options={"step_size": 1})
        (out * w).sum().backward()
        res["ref"] = (out.detach(), c.grad.clone())
        for p in func.parameters():
            p.grad = None
    fc = func.cuda()
    sl = slice(None) if rows is None else rows
    c = cref[sl].cuda().requires_grad_(True)
    X = tc.NaturalCubicSpline(c) if interp == "cubic" else tc.LinearInterpolation(c)
    out = tc.cdeint(X, fc, z0[sl].cuda(), X.grid_points, adjoint=False, method="rk4", options={"step_size": 1})
    (out * w[sl].cuda()).sum().backward()
    torch.cuda.synchronize()
    res["cuda"] = (out.detach(), c.grad.clone())
    return res


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(96, 12, 32, 64, 64, 3, "linear"), (67, 9, 7, 19, 23, 2, "cubic"),
                                   (40, 8, 100, 128, 128, 3, "linear")])
def test_path_gradient_against_oracle(tc, shape):
    """Config-shaped links of a stacked model (second link: C = hidden width of the first) and a cfg-5-shaped one."""
    B, K, C, H, HH, n, interp = shape
    res = _oracle_pair(tc, B, K, C, H, HH, n, interp, seed=3)
    assert rel(res["cuda"][0], res["ref"][0]) <= TOL
    assert rel(res["cuda"][1], res["ref"][1]) <= TOL


@pytest.mark.gpu
def test_path_gradient_full_size_properties(tc):
    """cfg-5 batch (1024 series, 100 channels, hidden 128): (1) dX/dt of a linear path depends on differences of
    consecutive knots only, so the gradient summed over the knots of any (series, channel) vanishes; (2) series are
    independent: a sub-batch solved alone gives the same path gradients (to rounding: the batch tiling may differ)."""
    full = _oracle_pair(tc, 1024, 6, 100, 128, 128, 3, "linear", seed=5, rows=slice(0, 1024))["cuda"][1]
    assert torch.isfinite(full).all()
    scale = float(full.abs().max())
    assert scale > 0
    assert float(full.sum(dim=1).abs().max()) <= 1e-5 * scale
    sub = _oracle_pair(tc, 1024, 6, 100, 128, 128, 3, "linear", seed=5, rows=slice(128, 320))["cuda"][1]
    assert rel(sub, full[128:320]) <= 1e-6


@pytest.mark.gpu
def test_path_gradient_unsupported_paths_are_loud(tc, gold):
    rec = gold["cdeint"]["lin_rk4_grid"]
    with pytest.raises(NotImplementedError):
        _run_cuda(tc, rec, precision="bf16")
    # adjoint=True without listing the coefficients: the reference warns and sends no gradient into them (solver.py:201-221)
    func = _field(rec).cuda()
    coeffs = rec["coeffs"].cuda().requires_grad_(True)
    X = tc.LinearInterpolation(coeffs, rec["knots"].cuda())
    z0 = rec["z0"].cuda().requires_grad_(True)
    with warnings.catch_warnings(record=True) as rec_w:
        warnings.simplefilter("always")
        out = tc.cdeint(X, func, z0, rec["t"].cuda(), adjoint=True, method="rk4", options={"step_size": 1.0})
    assert any("adjoint_params" in str(w.message) for w in rec_w)
    out.sum().backward()
    assert coeffs.grad is None and z0.grad is not None
    with pytest.raises(NotImplementedError):
        tc.cdeint(X, func, z0, rec["t"].cuda(), adjoint=True, method="rk4", options={"step_size": 1.0},
                  adjoint_params=tuple(func.parameters()) + (coeffs,))
