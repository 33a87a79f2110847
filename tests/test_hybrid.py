"""Linear / rectilinear hybrid path (src/ncde/interpolation.py:186-253).  Copy / select work: bit-exact.
Pinned by the reference's own known-answer test (src/tests/test_interpolation.py:6-34) and by vectors minted from the real
reference (tests/golden/make_hybrid_golden.py)."""
import os

import pytest
import torch

from oracle import cde_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hybrid.pt")


def _hand_case():
    nan = float("nan")
    times = torch.tensor([0.0, 1.0, 2.0, 3.0, 4.0])
    fast_data = torch.tensor([3.0, 1.4, nan, 3.4, nan])
    sparse_1 = torch.tensor([nan, 1.5, nan, nan, nan])
    sparse_2 = torch.tensor([nan, nan, nan, nan, 1.2])
    sparse_3 = torch.tensor([nan, nan, nan, nan, nan])
    data = torch.stack([times, fast_data, sparse_1, sparse_2, sparse_3]).T.unsqueeze(0)
    expect = torch.tensor([[[0.0000, 3.0000, 0.0000, 0.0000, 0.0000],
                            [1.0000, 1.4000, 0.0000, 0.0000, 0.0000],
                            [1.0000, 1.4000, 1.5000, 0.0000, 0.0000],
                            [2.0000, 2.4000, 1.5000, 0.0000, 0.0000],
                            [3.0000, 3.4000, 1.5000, 0.0000, 0.0000],
                            [4.0000, 3.4000, 1.5000, 0.0000, 0.0000],
                            [4.0000, 3.4000, 1.5000, 1.2000, 0.0000]]])
    return data, expect


def test_oracle_hand_case_of_the_reference():
    data, expect = _hand_case()
    assert torch.equal(O.prepare_linear_rectilinear_hybrid(data, [2, 3, 4]), expect)


def test_oracle_matches_vectors_from_the_real_reference():
    for case in torch.load(GOLDEN):
        assert torch.equal(O.prepare_linear_rectilinear_hybrid(case["x"].clone(), case["rect"]), case["out"])


@pytest.mark.gpu
def test_gpu_hand_case_of_the_reference():
    from ncde_b200 import interpolation
    data, expect = _hand_case()
    d = data.cuda()
    out = interpolation._prepare_linear_rectilinear_hybrid(d, rectilinear_indices=[2, 3, 4])
    assert torch.equal(out.cpu(), expect)
    # the input is mutated like the reference's: linear channel filled, first-row NaNs zeroed
    assert float(d[0, 2, 1]) == pytest.approx(2.4) and float(d[0, 0, 2]) == 0.0


@pytest.mark.gpu
def test_gpu_matches_vectors_from_the_real_reference_and_the_oracle():
    from ncde_b200 import interpolation
    for case in torch.load(GOLDEN):
        out = interpolation._prepare_linear_rectilinear_hybrid(case["x"].clone().cuda(), rectilinear_indices=case["rect"])
        assert torch.equal(out.cpu(), case["out"])
    # cfg-5 shaped: 100 channels, 72 steps, 20 regularly sampled channels
    torch.manual_seed(2)
    B, L, C = 64, 72, 100
    x = torch.randn(B, L, C)
    x[..., 0] = torch.arange(L, dtype=torch.float32)
    miss = torch.rand(B, L, C) > 0.1
    miss[..., :21] = torch.rand(B, L, 21) > 0.9
    miss[..., 0] = False
    x[miss] = float("nan")
    rect = list(range(21, C))
    ref = O.prepare_linear_rectilinear_hybrid(x.clone(), rect)
    out = interpolation._prepare_linear_rectilinear_hybrid(x.clone().cuda(), rectilinear_indices=rect)
    assert torch.equal(out.cpu(), ref)
    assert out.shape[1] <= 2 * L - 1
