"""Pins oracle/cde_oracle.py against vectors produced by the real reference (tests/golden/make_golden.py)."""
import warnings

import pytest
import torch

from oracle import cde_oracle as O


def same(a, b):
    """bit-equality treating NaN == NaN"""
    return a.shape == b.shape and a.dtype == b.dtype and torch.equal(torch.nan_to_num(a, nan=1234.5),
                                                                      torch.nan_to_num(b, nan=1234.5)) \
        and torch.equal(torch.isnan(a), torch.isnan(b))


def test_rectilinear_hand_case(golden_interp):
    # modules/torchcde/test/test_linear_interpolation.py:124-145
    rec = golden_interp["rect_hand"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = O.linear_interpolation_coeffs(rec["x"].clone(), rectilinear=rec["time_index"])
    assert same(got, rec["out"])
    x1_true = torch.tensor([[0.1, 0.2, 0.2, 0.9, 0.9], [0.4, 0.4, 0.4, 0.4, 1.1]]).T
    assert torch.equal(got[0], x1_true)


def test_rectilinear_random_exact(golden_interp):
    for rec in golden_interp["rect_random"]:
        x = rec["x"]
        assert same(O.forward_fill(x.clone()), rec["ffill"])
        assert same(O.rectilinear_prepare(x.clone(), rec["time_index"]), rec["rect_raw"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert same(O.linear_interpolation_coeffs(x.clone(), rectilinear=rec["time_index"]), rec["rect"])
            x0 = x.clone()
            got = O.linear_interpolation_coeffs(x0, rectilinear=rec["time_index"], initial_value_if_nan=0.25)
        assert same(got, rec["rect_init"])
        assert same(x0, rec["x_after_init"])  # in-place mutation is part of the contract


def test_rectilinear_nan_time_asserts(golden_interp):
    x = golden_interp["rect_hand"]["x"].clone()
    x[0, 1, 0] = float("nan")
    with pytest.raises(AssertionError):
        O.linear_interpolation_coeffs(x, rectilinear=0)


def test_linear_coeffs_nan_fill(golden_interp):
    for rec in golden_interp["linear_coeffs"]:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            got = O.linear_interpolation_coeffs(rec["x"].clone(), rec["t"])
            got_ff = O.linear_interpolation_coeffs(rec["x"].clone(), rec["t"], forward_fill_=True)
        assert same(got, rec["coeffs"])
        assert same(got_ff, rec["coeffs_ffill"])


def test_cubic_coeffs(golden_interp):
    for rec in golden_interp["cubic_coeffs"]:
        got1 = O.natural_cubic_coeffs(rec["x"].clone(), rec["t"], _version=1)
        got0 = O.natural_cubic_coeffs(rec["x"].clone(), rec["t"], _version=0)
        assert same(got1, rec["coeffs_v1"])
        assert same(got0, rec["coeffs_v0"])


def test_evaluate_derivative_index(golden_interp):
    for rec in golden_interp["evaluate"]:
        LX = O.LinearPath(rec["lin_coeffs"], rec["t"])
        CX = O.CubicPath(rec["cub_coeffs"], rec["t"])
        probes = rec["probes"]
        assert torch.equal(torch.stack([LX._locate(p)[1] for p in probes]), rec["lin_index"])
        assert torch.equal(torch.stack([CX._locate(p)[1] for p in probes]), rec["cub_index"])
        assert same(torch.stack([LX.evaluate(p) for p in probes], -2), rec["lin_eval"])
        assert same(torch.stack([LX.derivative(p) for p in probes], -2), rec["lin_deriv"])
        assert same(torch.stack([CX.evaluate(p) for p in probes], -2), rec["cub_eval"])
        assert same(torch.stack([CX.derivative(p) for p in probes], -2), rec["cub_deriv"])
        assert same(LX.evaluate(probes), rec["lin_eval_vec"])
        assert same(CX.derivative(probes), rec["cub_deriv_vec"])


def test_validate_errors():
    with pytest.raises(ValueError):
        O.linear_interpolation_coeffs(torch.zeros(3, 4, dtype=torch.int64))
    with pytest.raises(ValueError):
        O.linear_interpolation_coeffs(torch.zeros(4))
    with pytest.raises(ValueError):
        O.linear_interpolation_coeffs(torch.zeros(2, 1, 3))
    with pytest.raises(ValueError):
        O.natural_cubic_coeffs(torch.zeros(2, 4, 3), torch.tensor([0., 1., 1., 2.]))
    with pytest.raises(ValueError):
        O.CubicPath(torch.zeros(2, 4, 7))


def _field(rec):
    d = rec["dims"]
    if rec["field"] == "orig":
        f = O.SharedMLPField(d["C"], d["H"], d["HH"], d["n"])
    else:
        f = O.ToyField(d["C"], d["H"], width=d["HH"])
    f.load_state_dict(rec["state_dict"])
    return f


def run_oracle(rec, stats=None):
    func = _field(rec)
    X = O.CubicPath(rec["coeffs"]) if rec["interp"] == "cubic" else O.LinearPath(rec["coeffs"])
    z0 = rec["z0"].clone().requires_grad_(True)
    kw = rec["kw"]
    out = O.cdeint(X, func, z0, rec["t"], adjoint=kw["adjoint"], method=kw["method"], rtol=kw["rtol"],
                   atol=kw["atol"], options=kw["options"], stats=stats)
    (out * rec["w"]).sum().backward()
    grads = {n: p.grad for n, p in func.named_parameters()}
    return out.detach(), z0.grad, grads, func


FIXED = ["c1_toy_rect_rk4", "c2_lin_rk4_term", "c2_rect_rk4_online", "c2_lin_euler", "c5_small_rect",
         "cub_rk4_halfstep_offgrid", "lin_rk4_adjoint"]
ADAPTIVE = ["c3_cub_dopri5", "c3_cub_dopri5_adjoint", "cub_dopri5_free_online", "cub_dopri5_free_adjoint_online"]


def _close(a, b, tol):
    scale = b.abs().max().clamp_min(1e-30)
    return float((a - b).abs().max() / scale) <= tol


@pytest.mark.parametrize("name", FIXED)
def test_cdeint_fixed(golden_cdeint, name):
    rec = golden_cdeint[name]
    out, gz0, grads, func = run_oracle(rec)
    # same ATen ops in the same order: agreement is at rounding level (tolerance only guards BLAS blocking
    # differences between hosts)
    assert _close(out, rec["out"], 2e-6), name
    assert _close(gz0, rec["grad_z0"], 5e-6)
    for n, gref in rec["grads"].items():
        assert _close(grads[n], gref, 5e-6), n
    if rec["nfe"] is not None and not rec["kw"]["adjoint"]:
        assert func.nfe == rec["nfe"]


@pytest.mark.parametrize("name", ADAPTIVE)
def test_cdeint_adaptive(golden_cdeint, name):
    rec = golden_cdeint[name]
    stats = {}
    out, gz0, grads, func = run_oracle(rec, stats)
    assert func.nfe == rec["nfe"], (func.nfe, rec["nfe"], stats)  # identical accept/reject sequence
    assert _close(out, rec["out"], 5e-6), name
    assert _close(gz0, rec["grad_z0"], 2e-5)
    for n, gref in rec["grads"].items():
        assert _close(grads[n], gref, 2e-5), n
