"""vector_field_type 'evaluate' / 'derivative' (SURVEY §8f-2): dz/dt = f([z, X(t)]) or f([z, dX/dt(t)])
(modules/torchcde/torchcde/solver.py:112-137; src/ncde/vector_fields/base.py:56-104).

Golden vectors: tests/golden/vftype.pt from the REAL reference (tests/golden/make_vftype_golden.py).  CPU tests pin the oracle;
GPU tests compare the CUDA path (C ABI: ncde_problem_t.vf_type) with the golden vectors and with the oracle on config-shaped
problems.  Tolerance: relative max-norm 1e-5 (fp32).
"""
import os

import pytest
import torch

from oracle import cde_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vftype.pt")
TOL = 1e-5
CASES = ["eval_lin_rk4_grid", "deriv_lin_rk4_grid", "eval_cub_rk4_half_offgrid", "deriv_cub_euler_interval",
         "eval_rect_rk4_grid", "deriv_lin_rk4_nolayers", "eval_lin_rk4_wide"]


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLDEN)


def rel(a, b):
    a = a.detach().cpu()
    b = b.detach().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _field(rec):
    d = rec["dims"]
    f = O.SharedMLPField(d["C"], d["H"], d["HH"], d["n"], vector_field_type=rec["vector_field_type"])
    f.load_state_dict(rec["state_dict"])
    return f


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference(gold, name):
    rec = gold[name]
    func = _field(rec)
    X = O.CubicPath(rec["coeffs"]) if rec["interp"] == "cubic" else O.LinearPath(rec["coeffs"])
    z0 = rec["z0"].clone().requires_grad_(True)
    out = O.cdeint(X, func, z0, rec["t"], adjoint=False, method=rec["method"], options=dict(rec["options"]),
                   vector_field_type=rec["vector_field_type"])
    (out * rec["w"]).sum().backward()
    assert rel(out, rec["out"]) <= 1e-6
    assert rel(z0.grad, rec["grad_z0"]) <= 1e-5
    for n, p in func.named_parameters():
        assert rel(p.grad, rec["grads"][n]) <= 1e-5, n
    assert func.nfe == rec["nfe"]


def test_field_module_matches_reference_layout(gold):
    """ncde_b200.OriginalVectorField built for evaluate / derivative loads the reference's state_dict."""
    import ncde_b200
    for name in CASES:
        rec = gold[name]
        d = rec["dims"]
        f = ncde_b200.OriginalVectorField(d["C"], d["H"], d["HH"], d["n"], vector_field_type=rec["vector_field_type"])
        missing, unexpected = f.load_state_dict(rec["state_dict"])
        assert not missing and not unexpected
        assert f(None, torch.zeros(2, d["H"] + d["C"])).shape == (2, d["H"])


@pytest.fixture(scope="module")
def tc():
    import torchcde_b200
    assert torch.cuda.is_available()
    return torchcde_b200


def _run_cuda(tc, rec, **kw):
    import ncde_b200
    d = rec["dims"]
    func = ncde_b200.OriginalVectorField(d["C"], d["H"], d["HH"], d["n"], vector_field_type=rec["vector_field_type"])
    func.load_state_dict(rec["state_dict"])
    func = func.cuda()
    coeffs = rec["coeffs"].cuda()
    X = tc.NaturalCubicSpline(coeffs) if rec["interp"] == "cubic" else tc.LinearInterpolation(coeffs)
    z0 = rec["z0"].cuda().requires_grad_(True)
    args = dict(adjoint=False, vector_field_type=rec["vector_field_type"], method=rec["method"], options=dict(rec["options"]))
    args.update(kw)
    out = tc.cdeint(X, func, z0, rec["t"].cuda(), **args)
    (out * rec["w"].cuda()).sum().backward()
    torch.cuda.synchronize()
    return out, z0.grad, {n: p.grad for n, p in func.named_parameters()}, func


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_golden_vector_field_type(tc, gold, name):
    rec = gold[name]
    out, gz0, grads, func = _run_cuda(tc, rec)
    assert out.shape == rec["out"].shape
    assert rel(out, rec["out"]) <= TOL
    assert rel(gz0, rec["grad_z0"]) <= TOL
    for n, g in rec["grads"].items():
        assert rel(grads[n], g) <= TOL, n
    assert func.nfe == rec["nfe"]


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [("evaluate", 96, 20, 4, 64, 64, 3, "linear"), ("derivative", 67, 9, 7, 19, 23, 2, "cubic"),
                                   ("derivative", 72, 12, 100, 128, 128, 3, "linear")])
def test_vector_field_type_against_oracle(tc, shape):
    """cfg-2-, odd- and cfg-5-shaped problems (the `sparsity` ablation of the reference runs these modes on its configs)."""
    vft, B, K, C, H, HH, n, interp = shape
    g = torch.Generator().manual_seed(21)
    x = torch.randn(B, K, C, generator=g).cumsum(-2) * 0.2
    x[..., 0] = torch.arange(K, dtype=torch.float32)
    torch.manual_seed(4)
    func = O.SharedMLPField(C, H, HH, n, vector_field_type=vft)
    z0 = torch.randn(B, H, generator=g) * 0.5
    cref = O.natural_cubic_coeffs(x) if interp == "cubic" else x.clone()
    w = torch.randn(B, K, H, generator=g)
    Xr = O.CubicPath(cref) if interp == "cubic" else O.LinearPath(cref)
    z0r = z0.clone().requires_grad_(True)
    oref = O.cdeint(Xr, func, z0r, Xr.grid_points, adjoint=False, method="rk4", options={"step_size": 1}, vector_field_type=vft)
    (oref * w).sum().backward()
    gref = {k: p.grad.clone() for k, p in func.named_parameters()}
    for p in func.parameters():
        p.grad = None
    fc = func.cuda()
    c = cref.cuda()
    X = tc.NaturalCubicSpline(c) if interp == "cubic" else tc.LinearInterpolation(c)
    z0c = z0.cuda().requires_grad_(True)
    out = tc.cdeint(X, fc, z0c, X.grid_points, adjoint=False, vector_field_type=vft, method="rk4", options={"step_size": 1})
    (out * w.cuda()).sum().backward()
    torch.cuda.synchronize()
    assert rel(out, oref) <= TOL
    assert rel(z0c.grad, z0r.grad) <= TOL
    for k, p in fc.named_parameters():
        assert rel(p.grad, gref[k]) <= TOL, k


@pytest.mark.gpu
def test_vector_field_type_no_grad_forward_and_loud_errors(tc, gold):
    rec = gold["eval_lin_rk4_grid"]
    import ncde_b200
    d = rec["dims"]
    func = ncde_b200.OriginalVectorField(d["C"], d["H"], d["HH"], d["n"], vector_field_type="evaluate")
    func.load_state_dict(rec["state_dict"])
    func = func.cuda()
    X = tc.LinearInterpolation(rec["coeffs"].cuda())
    with torch.no_grad():   # forward-only path keeps one scratch record instead of the saved ones
        out = tc.cdeint(X, func, rec["z0"].cuda(), rec["t"].cuda(), adjoint=False, vector_field_type="evaluate", method="rk4",
                        options={"step_size": 1.0})
    assert rel(out, rec["out"]) <= TOL
    with pytest.raises(ValueError):
        tc.cdeint(X, func, rec["z0"].cuda(), rec["t"].cuda(), adjoint=False, vector_field_type="nonsense")
    with pytest.raises(ValueError):   # field built for another mode
        tc.cdeint(X, func, rec["z0"].cuda(), rec["t"].cuda(), adjoint=False, vector_field_type="derivative", method="rk4")
    for bad in (dict(method="dopri5"), dict(options={"step_size": 1.0, "precision": "bf16"})):
        with pytest.raises(NotImplementedError):
            _run_cuda(tc, rec, **bad)
