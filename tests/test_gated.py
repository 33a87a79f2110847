"""Gated vector fields: MinimalGatedVectorField — sigmoid(Linear_z(hh)) * tanh(Linear_r(hh)) — and
GRUGatedVectorField — sigmoid(Linear_z(net(h))) * tanh(Linear_r(net(sigmoid(Linear_f(h)) * h)))
(src/ncde/vector_fields/gating.py:7-61) — in all three vector_field_type modes (the reference's `sparsity` ablation,
experiments/configurations/configurations.json5: vector_field x vector_field_type, adjoint false).

Golden vectors: tests/golden/gated.pt from the REAL reference (tests/golden/make_gated_golden.py).  CPU tests pin the oracle and
the lowering; GPU tests compare the CUDA path (C ABI: ncde_mlp_t.W_gate) with the golden vectors and the oracle.
Tolerance: relative max-norm 1e-5 (fp32).
"""
import os

import pytest
import torch

from oracle import cde_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gated.pt")
TOL = 1e-5
CASES = ["min_matmul_lin_rk4", "min_matmul_cub_rk4_half", "min_matmul_rect_euler", "min_eval_lin_rk4", "min_deriv_cub_rk4",
         "min_matmul_wide", "gru_matmul_lin_rk4", "gru_matmul_cub_rk4_half", "gru_eval_rect_euler", "gru_deriv_lin_rk4"]


GPU_CASES = CASES


def _oracle_field(name):
    return O.GRUGatedField if name.startswith("gru") else O.MinimalGatedField


def _product_field(name):
    import ncde_b200
    return ncde_b200.GRUGatedVectorField if name.startswith("gru") else ncde_b200.MinimalGatedVectorField


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLDEN)


def rel(a, b):
    a = a.detach().cpu()
    b = b.detach().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference(gold, name):
    rec = gold[name]
    d = rec["dims"]
    func = _oracle_field(name)(d["C"], d["H"], d["HH"], d["n"], vector_field_type=rec["vector_field_type"])
    func.load_state_dict(rec["state_dict"])
    X = O.CubicPath(rec["coeffs"]) if rec["interp"] == "cubic" else O.LinearPath(rec["coeffs"])
    z0 = rec["z0"].clone().requires_grad_(True)
    out = O.cdeint(X, func, z0, rec["t"], adjoint=False, method=rec["method"], options=dict(rec["options"]),
                   vector_field_type=rec["vector_field_type"])
    (out * rec["w"]).sum().backward()
    assert rel(out, rec["out"]) <= 1e-6
    assert rel(z0.grad, rec["grad_z0"]) <= 1e-5
    for n, p in func.named_parameters():
        assert rel(p.grad, rec["grads"][n]) <= 1e-5, n


def test_lowering_of_gated_fields(gold):
    """MinimalGatedVectorField lowers to an MLP with a gate on its last layer; the GRU-gated field is refused loudly."""
    import ncde_b200
    from torchcde_b200 import lowering
    for name in [c for c in CASES if c.startswith("min")]:
        rec = gold[name]
        d = rec["dims"]
        f = ncde_b200.VECTOR_FIELDS["minimal"](d["C"], d["H"], d["HH"], d["n"], vector_field_type=rec["vector_field_type"])
        missing, unexpected = f.load_state_dict(rec["state_dict"])
        assert not missing and not unexpected
        spec = lowering.lower(f, d["H"], d["C"], rec["vector_field_type"])
        assert spec.gate is not None and spec.gate[0] is f.sigmoid_net[0].weight
        assert spec.weights[-1] is f.tanh_net[0].weight
        kinds = [(k, i) for _, k, i in spec.unique_params]
        assert kinds[-2:] == [("W", len(spec.weights)), ("b", len(spec.weights))]


@pytest.mark.parametrize("name", [c for c in CASES if c.startswith("gru")])
def test_gru_lowering_is_faithful_and_folds_gradients(gold, name):
    """The widened chain ([I; W_f] gate-in layer, block-diagonal hidden layers, K-stacked heads) equals the eager field, and
    autograd folds the gradients of the widened matrices back into the module's parameters (pure torch, no GPU)."""
    import ncde_b200
    from torchcde_b200 import _capi, lowering
    rec = gold[name]
    d = rec["dims"]
    vft = rec["vector_field_type"]
    f = ncde_b200.VECTOR_FIELDS["gru"](d["C"], d["H"], d["HH"], d["n"], vector_field_type=vft)
    missing, unexpected = f.load_state_dict(rec["state_dict"])
    assert not missing and not unexpected
    spec = lowering.lower(f, d["H"], d["C"], vft)
    assert spec.acts[0] == _capi.ACT_GATE_IN and spec.gate is not None
    assert spec.weights[0].shape == (2 * f.initial_dim, f.initial_dim)
    assert spec.weights[-1].shape == (f.output_dim, 2 * d["HH"])
    x = torch.randn(9, f.initial_dim, generator=torch.Generator().manual_seed(2))
    w = torch.randn(9, f.output_dim, generator=torch.Generator().manual_seed(3))
    want = f(None, x).reshape(9, -1)
    (want * w).sum().backward()
    gwant = {n: p.grad.clone() for n, p in f.named_parameters()}
    for p in f.parameters():
        p.grad = None
    got = spec.reference_forward(x, d["H"], d["C"] if vft == "matmul" else None).reshape(9, -1)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)
    (got * w).sum().backward()
    for n, p in f.named_parameters():
        assert torch.allclose(p.grad, gwant[n], rtol=1e-4, atol=1e-6), n


@pytest.fixture(scope="module")
def tc():
    import torchcde_b200
    assert torch.cuda.is_available()
    return torchcde_b200


def _run_cuda(tc, rec, **kw):
    d = rec["dims"]
    func = _product_field(rec["name"])(d["C"], d["H"], d["HH"], d["n"], vector_field_type=rec["vector_field_type"])
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
@pytest.mark.parametrize("name", GPU_CASES)
def test_golden_gated(tc, gold, name):
    rec = dict(gold[name], name=name)
    out, gz0, grads, func = _run_cuda(tc, rec)
    assert rel(out, rec["out"]) <= TOL
    assert rel(gz0, rec["grad_z0"]) <= TOL
    assert sorted(grads) == sorted(rec["grads"])
    for n, g in rec["grads"].items():
        assert rel(grads[n], g) <= TOL, n
    assert func.nfe == rec["nfe"]


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [("min", "matmul", 96, 20, 4, 64, 64, 3, "linear"), ("min", "matmul", 67, 9, 7, 19, 23, 2, "cubic"),
                                   ("min", "matmul", 80, 10, 64, 64, 48, 3, "linear"),
                                   ("min", "evaluate", 72, 12, 100, 128, 128, 3, "linear"),
                                   ("gru", "matmul", 96, 20, 4, 64, 64, 3, "linear"), ("gru", "matmul", 67, 9, 7, 19, 23, 2, "cubic"),
                                   ("gru", "derivative", 72, 10, 30, 64, 48, 3, "linear")])
def test_gated_against_oracle(tc, shape):
    kind, vft, B, K, C, H, HH, n, interp = shape
    g = torch.Generator().manual_seed(31)
    x = torch.randn(B, K, C, generator=g).cumsum(-2) * 0.2
    x[..., 0] = torch.arange(K, dtype=torch.float32)
    torch.manual_seed(6)
    func = _oracle_field(kind)(C, H, HH, n, vector_field_type=vft)
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
def test_gated_unsupported_combinations_are_loud(tc, gold):
    rec = dict(gold["min_matmul_lin_rk4"], name="min_matmul_lin_rk4")
    for bad in (dict(method="dopri5"), dict(options={"step_size": 1.0, "precision": "bf16"})):
        with pytest.raises(NotImplementedError):
            _run_cuda(tc, rec, **bad)
    import ncde_b200
    gru = ncde_b200.GRUGatedVectorField(3, 6, 160, 2).cuda()   # 2 * hidden_hidden_dim > 256: wider than the final-layer kernels take
    X = tc.LinearInterpolation(rec["coeffs"].cuda())
    with pytest.raises(NotImplementedError):
        tc.cdeint(X, gru, rec["z0"].cuda(), rec["t"].cuda(), adjoint=False, method="rk4", options={"step_size": 1.0})


# ---------------------------------------------------------------------------------------------------------------
# the same modes under the fixed-grid continuous adjoint (adjoint=True is NeuralCDE's default, src/ncde/ncde.py:60)
# ---------------------------------------------------------------------------------------------------------------
ADJ_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "adjoint_modes.pt")
ADJ_CASES = ["orig_eval_lin_rk4", "orig_deriv_cub_rk4_half", "min_matmul_lin_rk4", "min_eval_rect_euler", "gru_matmul_lin_rk4",
             "gru_deriv_lin_rk4"]


@pytest.fixture(scope="module")
def adj_gold():
    return torch.load(ADJ_GOLDEN)


def _adj_fields():
    import ncde_b200
    return {"orig": (O.SharedMLPField, ncde_b200.OriginalVectorField), "min": (O.MinimalGatedField, ncde_b200.MinimalGatedVectorField),
            "gru": (O.GRUGatedField, ncde_b200.GRUGatedVectorField)}


@pytest.mark.parametrize("name", ADJ_CASES)
def test_oracle_adjoint_modes_match_reference(adj_gold, name):
    rec = adj_gold[name]
    d = rec["dims"]
    func = _adj_fields()[rec["kind"]][0](d["C"], d["H"], d["HH"], d["n"], vector_field_type=rec["vector_field_type"])
    func.load_state_dict(rec["state_dict"])
    X = O.CubicPath(rec["coeffs"]) if rec["interp"] == "cubic" else O.LinearPath(rec["coeffs"])
    z0 = rec["z0"].clone().requires_grad_(True)
    out = O.cdeint(X, func, z0, rec["t"], adjoint=True, method=rec["method"], options=dict(rec["options"]),
                   vector_field_type=rec["vector_field_type"])
    (out * rec["w"]).sum().backward()
    assert rel(out, rec["out"]) <= 1e-6
    assert rel(z0.grad, rec["grad_z0"]) <= 1e-5
    for n, p in func.named_parameters():
        assert rel(p.grad, rec["grads"][n]) <= 1e-5, n


@pytest.mark.gpu
@pytest.mark.parametrize("name", ADJ_CASES)
def test_golden_adjoint_modes(tc, adj_gold, name):
    rec = adj_gold[name]
    d = rec["dims"]
    func = _adj_fields()[rec["kind"]][1](d["C"], d["H"], d["HH"], d["n"], vector_field_type=rec["vector_field_type"])
    func.load_state_dict(rec["state_dict"])
    func = func.cuda()
    coeffs = rec["coeffs"].cuda()
    X = tc.NaturalCubicSpline(coeffs) if rec["interp"] == "cubic" else tc.LinearInterpolation(coeffs)
    z0 = rec["z0"].cuda().requires_grad_(True)
    out = tc.cdeint(X, func, z0, rec["t"].cuda(), adjoint=True, vector_field_type=rec["vector_field_type"], method=rec["method"],
                    options=dict(rec["options"]))
    (out * rec["w"].cuda()).sum().backward()
    torch.cuda.synchronize()
    assert rel(out, rec["out"]) <= TOL
    assert rel(z0.grad, rec["grad_z0"]) <= TOL
    grads = {n: p.grad for n, p in func.named_parameters()}
    assert sorted(grads) == sorted(rec["grads"])
    for n, g in rec["grads"].items():
        assert rel(grads[n], g) <= TOL, n
