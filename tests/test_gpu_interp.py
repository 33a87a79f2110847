"""GPU parity of the interpolation constructors / evaluators against the golden vectors of the real reference and
against the oracle.  Bit-exact (torch.equal) everywhere: csrc/interp.cu is compiled without FMA contraction."""
import warnings

import pytest
import torch

from oracle import cde_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tc():
    import torchcde_b200
    assert torch.cuda.is_available()
    return torchcde_b200


def same(a, b):
    a = a.cpu()
    return a.shape == b.shape and a.dtype == b.dtype and torch.equal(torch.isnan(a), torch.isnan(b)) and \
        torch.equal(torch.nan_to_num(a, nan=7.0), torch.nan_to_num(b, nan=7.0))


def test_rectilinear_hand_case(tc, golden_interp):
    rec = golden_interp["rect_hand"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = tc.linear_interpolation_coeffs(rec["x"].cuda(), rectilinear=0)
        assert same(got, rec["out"])
        # 2-D, 4-D and swapped-time variants of modules/torchcde/test/test_linear_interpolation.py:137-151
        x = rec["x"]
        assert same(tc.linear_interpolation_coeffs(x[0].cuda(), rectilinear=0), rec["out"][0])
        x4 = torch.stack([x, x])
        assert same(tc.linear_interpolation_coeffs(x4.cuda(), rectilinear=0), torch.stack([rec["out"]] * 2))
        sw = tc.linear_interpolation_coeffs(x[:, :, [1, 0]].cuda(), rectilinear=1)
        assert same(sw, rec["out"][:, :, [1, 0]])
    bad = x.clone()
    bad[0, 1, 0] = float("nan")
    with pytest.raises(AssertionError):
        tc.linear_interpolation_coeffs(bad.cuda(), rectilinear=0)


def test_rectilinear_random(tc, golden_interp):
    for rec in golden_interp["rect_random"]:
        x = rec["x"]
        assert same(tc.forward_fill(x.cuda()), rec["ffill"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            from torchcde_b200.interpolation_linear import _prepare_rectilinear_interpolation
            assert same(_prepare_rectilinear_interpolation(x.cuda(), rec["time_index"]), rec["rect_raw"])
            assert same(tc.linear_interpolation_coeffs(x.cuda(), rectilinear=rec["time_index"]), rec["rect"])
            x0 = x.clone().cuda()
            got = tc.linear_interpolation_coeffs(x0, rectilinear=rec["time_index"], initial_value_if_nan=0.25)
        assert same(got, rec["rect_init"])
        assert same(x0, rec["x_after_init"])


def test_linear_fill(tc, golden_interp):
    for rec in golden_interp["linear_coeffs"]:
        t = None if rec["t"] is None else rec["t"].cuda()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            xin = rec["x"].cuda()
            keep = xin.clone()
            got = tc.linear_interpolation_coeffs(xin, t)
            assert same(xin, keep.cpu()), "input must not be modified on the NaN-fill path"
            got_ff = tc.linear_interpolation_coeffs(rec["x"].cuda(), t, forward_fill=True)
        assert same(got, rec["coeffs"])
        assert same(got_ff, rec["coeffs_ffill"])
    # nothing to do -> the very same object comes back (interpolation_linear.py:180)
    x = torch.randn(2, 5, 3).cuda()
    assert tc.linear_interpolation_coeffs(x) is x


def test_cubic_coeffs(tc, golden_interp):
    for rec in golden_interp["cubic_coeffs"]:
        t = None if rec["t"] is None else rec["t"].cuda()
        assert same(tc.natural_cubic_coeffs(rec["x"].cuda(), t), rec["coeffs_v1"])
        assert same(tc.natural_cubic_spline_coeffs(rec["x"].cuda(), t), rec["coeffs_v0"])


def test_evaluate_derivative_index(tc, golden_interp):
    for rec in golden_interp["evaluate"]:
        t = None if rec["t"] is None else rec["t"].cuda()
        LX = tc.LinearInterpolation(rec["lin_coeffs"].cuda(), t)
        CX = tc.NaturalCubicSpline(rec["cub_coeffs"].cuda(), t)
        probes = rec["probes"].cuda()
        assert torch.equal(LX.knot_index(probes).cpu(), rec["lin_index"])
        assert torch.equal(CX.knot_index(probes).cpu(), rec["cub_index"])
        assert same(LX.evaluate(probes), rec["lin_eval"])
        assert same(LX.derivative(probes), rec["lin_deriv"])
        assert same(CX.evaluate(probes), rec["cub_eval"])
        assert same(CX.derivative(probes), rec["cub_deriv"])
        # scalar query
        assert same(LX.evaluate(probes[3]), rec["lin_eval"][..., 3, :])
        assert same(CX.derivative(float(rec["probes"][5])), rec["cub_deriv"][..., 5, :])
        assert same(LX.interval, torch.stack([rec["probes"][0], LX.grid_points.cpu()[-1]]))


def test_large_random_against_oracle(tc):
    """Config-sized inputs (cfg 5: 1024 series x 72 steps x 100 channels, 80% missing) against the oracle's
    vectorised constructors; cubic on cfg-3 shape."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(256, 72, 100, generator=g)
    x[..., 0] = torch.arange(72.)
    drop = torch.rand(x.shape, generator=g) < 0.8
    drop[..., 0] = False
    drop[:, 0, :] = False
    x[drop] = float("nan")
    want = O.rectilinear_prepare(x.clone(), 0)
    got = tc.linear_interpolation_coeffs(x.cuda(), rectilinear=0)
    assert same(got, want)
    assert got.shape == (256, 143, 100)
    # idempotence: filled data has no NaN, so a second fill is the identity
    assert same(tc.forward_fill(got), want)
    xs = torch.randn(64, 161, 21, generator=g)
    assert same(tc.natural_cubic_coeffs(xs.cuda()), O.natural_cubic_coeffs(xs))
    xd = torch.randn(8, 40, 6, generator=g, dtype=torch.float64)
    xd[torch.rand(xd.shape, generator=g) < 0.3] = float("nan")
    assert same(tc.natural_cubic_coeffs(xd.cuda()), O.natural_cubic_coeffs(xd))
    assert same(tc.linear_interpolation_coeffs(xd.cuda()), O.linear_interpolation_coeffs(xd.clone()))


def test_validation_errors(tc):
    with pytest.raises(ValueError):
        tc.linear_interpolation_coeffs(torch.zeros(3, 4, dtype=torch.int64).cuda())
    with pytest.raises(ValueError):
        tc.linear_interpolation_coeffs(torch.zeros(4).cuda())
    with pytest.raises(ValueError):
        tc.linear_interpolation_coeffs(torch.zeros(2, 1, 3).cuda())
    with pytest.raises(ValueError):
        tc.natural_cubic_coeffs(torch.zeros(2, 4, 3).cuda(), torch.tensor([0., 1., 1., 2.]).cuda())
    with pytest.raises(ValueError):
        tc.NaturalCubicSpline(torch.zeros(2, 4, 7).cuda())
