"""Offline preprocessing over ragged series: one ncde_ragged_interpolate launch for all series against the
reference's per-series Python loop.

Golden vectors: tests/golden/ragged.pt — `transform` from the REAL get_data/transformers.py; `loader` restated from
experiments/ingredients/loader.py:100-113,181-202 (that file imports sacred / ignite / autots, absent here; see
tests/golden/make_ragged_golden.py).  Copy / select / count work and the same rounding per operation: BIT-EXACT (torch.equal)
for float32 and float64.
"""
import os

import pytest
import torch

from oracle import cde_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ragged.pt")
METHODS = ["linear", "rectilinear", "cubic", "linear_forward_fill"]


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLDEN)


def same(a, b):
    return a.shape == b.shape and a.dtype == b.dtype and torch.equal(torch.nan_to_num(a.cpu(), nan=1234.5), torch.nan_to_num(b, nan=1234.5))


# ---------------------------------------------------------------------------------------------------------------
# CPU: the oracle restatement against the real transformer
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["mimic_like", "dense", "f64"])
@pytest.mark.parametrize("method", METHODS)
def test_oracle_transform_matches_reference(gold, name, method):
    rec = gold["transform"][name]
    data = [x.clone() for x in rec["raw"]]
    out = O.interpolation_transform(data, method)
    assert len(out) == len(rec[method])
    for a, b in zip(out, rec[method]):
        assert same(a, b)
    for a, b in zip(data, rec["mutated_" + method]):   # the in-place side effect on the caller's series
        assert same(a, b)


def test_oracle_loader_pieces_match(gold):
    rec = gold["loader"]
    data = [x.clone() for x in rec["raw"]]
    rect = O.interpolation_transform(data, "rectilinear")
    with_int = [O.rectilinear_intensity(c, d) for c, d in zip(rect, data)]
    for a, b in zip(with_int, rec["with_intensity"]):
        assert same(a, b)
    batches = O.padded_batches([with_int[i] for i in rec["order"]], rec["batch_size"])
    for a, b in zip(batches, rec["batches"]):
        assert same(a, b)


def test_host_side_contract_without_a_gpu():
    from ncde_b200 import preprocessing as P
    with pytest.raises(AssertionError):
        P.Interpolation(method="nonsense")
    with pytest.raises(NotImplementedError):
        P.Interpolation(method="hybrid", channel_indices=[1])
    assert repr(P.Interpolation("rectilinear")) == "Rectilinear Interpolation"
    with pytest.raises(ValueError):   # a series of length 1 (torchcde/misc.py:70-100)
        P.ragged_interpolate([torch.zeros(1, 3), torch.zeros(4, 3)], "linear")
    with pytest.raises(RuntimeError):   # no CPU fallback
        if torch.cuda.is_available():
            raise RuntimeError("has a GPU")
        P.ragged_interpolate([torch.zeros(3, 3), torch.zeros(4, 3)], "linear", device=torch.device("cpu"))
    s, t, l, idx = P.sort_unequal_lengths(None, [torch.zeros(5, 2), torch.zeros(2, 2), torch.zeros(3, 2)], [0, 1, 2])
    assert idx == [1, 2, 0] and [len(x) for x in t] == [2, 3, 5] and l == [1, 2, 0]


# ---------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def P():
    from ncde_b200 import preprocessing
    assert torch.cuda.is_available()
    return preprocessing


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["mimic_like", "dense", "f64"])
@pytest.mark.parametrize("method", METHODS)
def test_golden_transform(P, gold, name, method):
    rec = gold["transform"][name]
    data = [x.clone() for x in rec["raw"]]            # CPU series in, CPU coefficients out, like the reference
    out = P.Interpolation(method=method).fit_transform(data)
    assert len(out) == len(rec[method])
    for a, b in zip(out, rec[method]):
        assert not a.is_cuda and same(a, b)
    for a, b in zip(data, rec["mutated_" + method]):
        assert same(a, b)
    dev = [x.clone().cuda() for x in rec["raw"]]      # series already on the GPU stay there
    out = P.Interpolation(method=method).fit_transform(dev)
    assert all(a.is_cuda and same(a, b) for a, b in zip(out, rec[method]))


@pytest.mark.gpu
def test_golden_transform_tensor_input(P, gold):
    rec = gold["transform"]["tensor"]
    for method in ("linear", "rectilinear", "cubic"):
        out = P.Interpolation(method=method).fit_transform(rec["raw"].clone().cuda())
        assert isinstance(out, torch.Tensor) and same(out, rec[method])


@pytest.mark.gpu
def test_golden_intensity_and_padded_batches(P, gold):
    rec = gold["loader"]
    coeffs, rows = P.rectilinear_intensity([x.clone() for x in rec["raw"]], pad=True)
    assert coeffs.shape[-1] == 2 * rec["raw"][0].shape[-1] - 1
    for i, want in enumerate(rec["with_intensity"]):
        assert rows[i] == want.shape[0] and same(coeffs[i, :rows[i]], want)
    # loader.py:158-166,181-202: sort by length, then batches padded to their own longest series by repeating last rows
    order = sorted(range(len(rows)), key=lambda k: rows[k])
    assert order == rec["order"]
    batches = P.padded_batches(coeffs[order], [rows[i] for i in order], rec["batch_size"])
    assert len(batches) == len(rec["batches"])
    for a, b in zip(batches, rec["batches"]):
        assert same(a, b)
    # pad=False leaves the rows past a series' end missing
    nopad, _ = P.rectilinear_intensity([x.clone() for x in rec["raw"]], pad=False)
    short = min(range(len(rows)), key=lambda k: rows[k])
    assert torch.isnan(nopad[short, rows[short]:]).all() and same(nopad[short, :rows[short]], rec["with_intensity"][short])


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["linear", "rectilinear", "cubic"])
def test_full_size_ragged_equals_per_series_kernels(P, method):
    """cfg-5-sized raw set (8192 series, up to 72 rows, 100 channels, ~75 % missing): a sample of series must be bit-identical
    to the fixed-length kernels run on that series alone, every valid row must be finite, and the time channel stays sorted."""
    import torchcde_b200 as tc
    g = torch.Generator().manual_seed(8)
    n, Lmax, C = 8192, 72, 100
    lengths = torch.randint(2, Lmax + 1, (n,), generator=g)
    x = torch.randn(n, Lmax, C, generator=g)
    x[..., 0] = torch.arange(Lmax, dtype=torch.float32)
    drop = torch.rand(n, Lmax, C, generator=g) < 0.75
    drop[..., 0] = False
    x[drop] = float("nan")
    series = [x[i, :int(lengths[i])].clone().cuda() for i in range(n)]
    coeffs, rows = P.ragged_interpolate(series, method, pad=False)
    torch.cuda.synchronize()
    valid = torch.arange(coeffs.shape[1], device="cuda")[None, :] < torch.tensor(rows, device="cuda")[:, None]
    assert torch.isfinite(coeffs[valid]).all()
    assert torch.isnan(coeffs[~valid]).all()
    if method != "cubic":
        t = coeffs[..., 0]
        step = (t[:, 1:] - t[:, :-1])[valid[:, 1:]]
        assert (step >= 0).all()
    for i in list(range(0, n, 911)) + [int(lengths.argmin()), int(lengths.argmax())]:
        d = series[i].clone()
        d[:1][torch.isnan(d[:1])] = 0.0
        if method == "cubic":
            want = tc.natural_cubic_coeffs(d)
        else:
            want = tc.linear_interpolation_coeffs(d, rectilinear=0 if method == "rectilinear" else None)
        assert torch.equal(coeffs[i, :rows[i]], want), i
