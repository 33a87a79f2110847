"""AttentionNeuralCDE (src/ncde/attention.py) against the REAL reference module: golden vectors
tests/golden/attention.pt (tests/golden/make_attention_golden.py runs the reference with `autots` stubbed): softmax and sparsemax
attention, static features, backprop through the solvers (gradients flow through the hidden-state control paths) and the
continuous adjoint, forwards and backwards attention.  Tolerance: relative max-norm 1e-5 (fp32)."""
import os
import warnings

import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "attention.pt")
TOL = 1e-5
CASES = ["softmax_backprop", "sparsemax_backprop_static", "softmax_adjoint_forwards"]


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLDEN)


def rel(a, b):
    a, b = a.detach().cpu(), b.detach().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_sparsemax_matches_the_reference_module(gold):
    from ncde_b200.attention import Sparsemax
    rec = gold["sparsemax"]
    out = Sparsemax(dim=1)(rec["in"])
    assert torch.allclose(out, rec["out"], atol=1e-6)
    assert torch.allclose(out.sum(1), torch.ones(4, 1), atol=1e-6) and (out >= 0).all() and (out == 0).any()


def test_keep_and_pad_is_pad_plus_forward_fill():
    from ncde_b200.attention import keep_and_pad
    torch.manual_seed(0)
    h = torch.randn(5, 7, 3, requires_grad=True)
    keep = torch.rand(5, 7) > 0.5
    keep[:, 2] = True
    out = keep_and_pad(h, keep)
    n = int(keep.sum(1).max())
    assert out.shape == (5, n, 3)
    for b in range(5):
        rows = h[b][keep[b]]
        expect = torch.cat([rows, rows[-1:].expand(n - rows.size(0), -1)])
        assert torch.equal(out[b], expect)
    out.sum().backward()
    assert h.grad is not None and float(h.grad[~keep].abs().max()) == 0.0
    with pytest.raises(ValueError):
        keep_and_pad(h, torch.zeros(5, 7, dtype=torch.bool))


@pytest.mark.parametrize("name", CASES)
def test_state_dict_is_interchangeable_with_the_reference(gold, name):
    import ncde_b200
    rec = gold[name]
    C, H, O_ = rec["dims"]
    m = ncde_b200.AttentionNeuralCDE(C, H, O_, **rec["kwargs"])
    missing, unexpected = m.load_state_dict(rec["state_dict"])
    assert not missing and not unexpected


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_golden_attention(gold, name):
    import ncde_b200
    rec = gold[name]
    C, H, O_ = rec["dims"]
    m = ncde_b200.AttentionNeuralCDE(C, H, O_, **rec["kwargs"])
    m.load_state_dict(rec["state_dict"])
    m = m.cuda()
    coeffs = rec["coeffs"].cuda()
    inputs = coeffs if rec["static"] is None else [rec["static"].cuda(), coeffs]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        hidden = m.encoder(inputs)
        att = m.attention(hidden if rec["static"] is None else [rec["static"].cuda(), hidden])
        assert rel(att, rec["attention"]) <= TOL
        assert torch.equal((att > 1.0 / hidden.size(1)).sum(1).flatten().cpu(), rec["kept"])
        for p in m.parameters():
            p.grad = None
        y = m(inputs)
    assert y.shape == rec["out"].shape
    (y * rec["w"].cuda()).sum().backward()
    torch.cuda.synchronize()
    assert rel(y, rec["out"]) <= TOL
    got = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    assert sorted(got) == sorted(rec["grads"])
    for k, g in rec["grads"].items():
        assert rel(got[k], g) <= 2 * TOL, (k, rel(got[k], g))
