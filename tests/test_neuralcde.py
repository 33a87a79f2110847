"""The model wrapper (h0 = Linear([static ⊕] X(0)), cdeint, readout) end to end against the REAL reference
``src/ncde/ncde.py`` — golden vectors tests/golden/neuralcde.pt (tests/golden/make_neuralcde_golden.py): vector fields original /
minimal / gru, vector_field_type matmul / evaluate / derivative, linear / rectilinear / cubic / smoothed paths, static features,
online and terminal outputs, backprop through the solver and the continuous adjoint.  Tolerance: relative max-norm 1e-5 (fp32).
"""
import os
import warnings

import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "neuralcde.pt")
TOL = 1e-5
CASES = ["orig_rect_static_online", "orig_cubic_adjoint_terminal", "orig_linear_noinitial", "minimal_evaluate_online",
         "gru_derivative_static_terminal", "gru_matmul_rect_online", "orig_smooth_cubic"]


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLDEN)


def rel(a, b):
    a = a.detach().cpu()
    b = b.detach().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _model(rec):
    import ncde_b200
    C, H, O_ = rec["dims"]
    m = ncde_b200.NeuralCDE(C, H, O_, **rec["kwargs"])
    missing, unexpected = m.load_state_dict(rec["state_dict"])
    assert not missing and not unexpected
    return m


@pytest.mark.parametrize("name", CASES)
def test_state_dict_is_interchangeable_with_the_reference(gold, name):
    m = _model(gold[name])
    assert sorted(k for k, _ in m.named_parameters()) == sorted(gold[name]["grads"]) or gold[name]["kwargs"].get("adjoint")


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_golden_neuralcde(gold, name):
    rec = gold[name]
    m = _model(rec).cuda()
    coeffs = rec["coeffs"].cuda()
    inputs = coeffs if rec["static"] is None else (rec["static"].cuda(), coeffs)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        y = m(inputs)
    assert y.shape == rec["out"].shape
    (y * rec["w"].cuda()).sum().backward()
    torch.cuda.synchronize()
    assert rel(y, rec["out"]) <= TOL
    got = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    assert sorted(got) == sorted(rec["grads"])
    for k, g in rec["grads"].items():
        assert rel(got[k], g) <= TOL, k
    if not rec["kwargs"].get("adjoint"):
        assert m.nfe == rec["nfe"]
