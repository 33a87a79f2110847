"""Decreasing integration times and options['perturb'] (torchdiffeq _check_inputs / _PerturbFunc,
modules/torchdiffeq/torchdiffeq/_impl/misc.py:168-191,262-283) on the fixed-grid solvers, against golden vectors minted by the
REAL reference (tests/golden/make_reverse_golden.py).  Tolerance: relative max-norm 1e-5 (fp32)."""
import os

import numpy as np
import pytest
import torch

from oracle import cde_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reverse_perturb.pt")
CASES = ["rev_lin_rk4_grid", "rev_cub_rk4_half_offgrid", "rev_rect_euler_grid", "perturb_rect_rk4_grid", "perturb_lin_euler_grid",
         "rev_perturb_lin_rk4"]
TOL = 1e-5


def rel(a, b):
    a, b = a.detach().cpu(), b.detach().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_schedule_reversed_and_perturbed():
    from torchcde_b200.solver import FixedSchedule
    t = torch.tensor([0., 1., 2.5, 4.])
    fwd = FixedSchedule(t, "rk4", 0.5, None, None, None)
    rev = FixedSchedule(t.flip(0), "rk4", 0.5, None, None, None)
    assert rev.reverse and not fwd.reverse and rev.n_steps == fwd.n_steps == 8
    # reversed: stage times run downwards from t[-1] = 4, dt is negative, the first stage of the first step is the start time
    assert rev.stage_t[0, 0] == 4.0 and rev.stage_t[-1, -1] == 0.0 and (rev.dt < 0).all() and (fwd.dt > 0).all()
    assert np.all(np.diff(rev.stage_t.reshape(-1)) <= 0)
    p = FixedSchedule(t, "rk4", 0.5, None, None, None, perturb=True)
    assert np.array_equal(p.stage_t[:, 0], np.nextafter(fwd.stage_t[:, 0], np.float32(np.inf)))
    assert np.array_equal(p.stage_t[:, 3], np.nextafter(fwd.stage_t[:, 3], np.float32(-np.inf)))
    assert np.array_equal(p.stage_t[:, 1:3], fwd.stage_t[:, 1:3])
    e = FixedSchedule(t, "euler", 0.5, None, None, None, perturb=True)
    assert e.stage_t.shape == (8, 1) and np.array_equal(e.stage_t[:, 0], np.nextafter(fwd.stage_t[:, 0], np.float32(np.inf)))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_golden_reverse_perturb(name):
    import torchcde_b200 as tc
    rec = torch.load(GOLDEN)[name]
    d = rec["dims"]
    func = O.SharedMLPField(d["C"], d["H"], d["HH"], d["n"])
    func.load_state_dict(rec["state_dict"])
    func = func.cuda()
    coeffs = rec["coeffs"].cuda()
    X = tc.NaturalCubicSpline(coeffs) if rec["interp"] == "cubic" else tc.LinearInterpolation(coeffs)
    z0 = rec["z0"].cuda().requires_grad_(True)
    kw = rec["kw"]
    out = tc.cdeint(X, func, z0, rec["t"].cuda(), adjoint=False, method=kw["method"], rtol=kw["rtol"], atol=kw["atol"],
                    options=dict(kw["options"]))
    (out * rec["w"].cuda()).sum().backward()
    torch.cuda.synchronize()
    assert out.shape == rec["out"].shape
    assert rel(out, rec["out"]) <= TOL
    assert rel(z0.grad, rec["grad_z0"]) <= TOL
    for n, g in rec["grads"].items():
        assert rel(dict(func.named_parameters())[n].grad, g) <= TOL, n
    assert func.nfe == rec["nfe"]


@pytest.mark.gpu
def test_reversed_time_on_the_tensor_core_path():
    """The persistent kernels take the (negative) step sizes and decreasing stage times from the same schedule."""
    import copy
    import torchcde_b200 as tc
    rec = torch.load(GOLDEN)["rev_lin_rk4_grid"]
    d = rec["dims"]
    func = O.SharedMLPField(d["C"], d["H"], d["HH"], d["n"])
    func.load_state_dict(rec["state_dict"])
    X = tc.LinearInterpolation(rec["coeffs"].cuda())
    for prec, tol in (("bf16x3", 2e-4), ("bf16", 3e-2)):
        f = copy.deepcopy(func).cuda()
        z0 = rec["z0"].cuda().requires_grad_(True)
        out = tc.cdeint(X, f, z0, rec["t"].cuda(), adjoint=False, method="rk4", options={"step_size": 1, "precision": prec})
        (out * rec["w"].cuda()).sum().backward()
        assert rel(out, rec["out"]) <= tol, (prec, rel(out, rec["out"]))
        assert rel(z0.grad, rec["grad_z0"]) <= 10 * tol
