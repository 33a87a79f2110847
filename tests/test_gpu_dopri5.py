"""dopri5 with device-side step control (csrc/adaptive_kernels.cuh) against the oracle and the reference's goldens.

What can be pinned exactly is pinned exactly: the RK step, first-same-as-last, the dense output, the output
bookkeeping (forced step sequences, 1e-6), the initial step selection (to fp32 rounding) and the whole accept/reject
sequence whenever the error estimates are above rounding noise (identical NFE).  With automatic control the very first
error estimate of a tiny initial step is pure fp32 cancellation noise (k . c_err with sum(c_err) = 0), so two correct
implementations grow that step by different factors and then follow different, equally valid, step sequences: there
the check is the global error against a tight-tolerance solution, which must be as small as the reference's own."""
import pytest
import torch

from oracle import cde_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def problem():
    import torchcde_b200 as tc
    torch.manual_seed(0)
    x = torch.rand(3, 8, 2)
    func = O.SharedMLPField(2, 4, 8, 1)
    z0 = torch.rand(3, 4)
    Xr = O.CubicPath(O.natural_cubic_coeffs(x))
    X = tc.NaturalCubicSpline(tc.natural_cubic_coeffs(x.cuda()))
    fd = O.SharedMLPField(2, 4, 8, 1).cuda()
    fd.load_state_dict(func.state_dict())
    return Xr, func, X, fd, z0


def both(problem, t, **kw):
    import torchcde_b200 as tc
    from torchcde_b200 import adaptive
    Xr, func, X, fd, z0 = problem
    t = torch.tensor(t, dtype=torch.float64)
    st = {}
    with torch.no_grad():
        ref = O.cdeint(Xr, func, z0, t, adjoint=False, method="dopri5", stats=st, **kw)
        out = tc.cdeint(X, fd, z0.cuda(), t.cuda(), adjoint=False, method="dopri5", **kw)
    return ref, out, st, dict(adaptive.last_stats)


def test_forced_sequences_are_exact(problem):
    """rtol = atol = 1 accepts every step: pins the Dormand-Prince step, FSAL, dense output and output bookkeeping."""
    for t, opts in [([0., 0.05, 0.1, 0.2, 0.3], {"first_step": 0.3}),
                    ([0., 0.3, 0.9], {"first_step": 0.3, "max_step": 0.3}),
                    ([0., 2.0, 6.5], {"first_step": 0.3, "max_step": 0.3}),
                    ([0., 7.0], {"first_step": 0.25, "max_step": 0.5, "min_step": 0.5})]:
        ref, out, st, gst = both(problem, t, rtol=1.0, atol=1.0, options=opts)
        assert (gst["attempted"], gst["accepted"], gst["nfe"]) == (st["attempted"], st["accepted"], st["nfe"])
        assert rel(out, ref) <= 2e-6


def test_control_sequence_matches_with_given_first_step(problem):
    """Real tolerances, error estimates above rounding noise: identical accept / reject sequence."""
    for t, kw in [([0., 2.0], dict(rtol=1e-4, atol=1e-6, options={"first_step": 0.05})),
                  ([0., 7.0], dict(rtol=1e-3, atol=1e-5, options={"first_step": 0.05})),
                  ([0., 3.0, 7.0], dict(rtol=1e-3, atol=1e-5, options={"first_step": 0.05, "min_step": 0.5}))]:
        ref, out, st, gst = both(problem, t, **kw)
        assert (gst["attempted"], gst["accepted"], gst["nfe"]) == (st["attempted"], st["accepted"], st["nfe"])
        assert rel(out, ref) <= 5e-4      # differences of 1e-7 per step are amplified by the step-size feedback


def test_initial_step_selection_matches(problem):
    import oracle.cde_oracle as OO
    seen = {}
    orig = OO._initial_step

    def spy(*a, **k):
        seen["h"] = float(orig(*a, **k))
        return orig(*a, **k)
    OO._initial_step = spy
    try:
        ref, out, st, gst = both(problem, [0., 2.0], rtol=1e-4, atol=1e-6, options={})
    finally:
        OO._initial_step = orig
    assert abs(gst["first_step"] - seen["h"]) <= 4e-7 * seen["h"]      # misc.py:32-71, fp32 arithmetic


def test_automatic_control_global_error(problem):
    """Global error against a tight solution is as small as the reference's own (within 3x)."""
    for rtol, atol, t in [(1e-4, 1e-6, [0., 2.0, 7.0]), (1e-3, 1e-5, [0., 7.0])]:
        ref, out, st, gst = both(problem, t, rtol=rtol, atol=atol, options={})
        Xr, func, X, fd, z0 = problem
        with torch.no_grad():
            truth = O.cdeint(Xr.__class__(torch.cat([Xr.a, Xr.b, Xr.two_c, Xr.three_d], -1).double()), func.double(),
                             z0.double(), torch.tensor(t, dtype=torch.float64), adjoint=False, method="dopri5",
                             rtol=1e-10, atol=1e-12)
            func.float()
        e_ref, e_gpu = rel(ref, truth), rel(out, truth)
        assert e_gpu <= 3 * e_ref + 1e-6, (e_gpu, e_ref)
        assert abs(gst["nfe"] - st["nfe"]) <= 0.35 * st["nfe"]


@pytest.mark.parametrize("name", ["c3_cub_dopri5", "cub_dopri5_free_online"])
def test_golden_forward(golden_cdeint, name):
    """Vectors from the real reference.  The dense output between step ends is only 4th order and the steps span
    spline knots, so the reference's own outputs are off by up to 1e-2 at interior times; the bound is therefore the
    reference's own error against a tight fp64 solution (x3), plus NFE in the same range."""
    import torchcde_b200 as tc
    rec = golden_cdeint[name]
    d = rec["dims"]
    func = O.SharedMLPField(d["C"], d["H"], d["HH"], d["n"])
    func.load_state_dict(rec["state_dict"])
    kw = rec["kw"]
    with torch.no_grad():
        truth = O.cdeint(O.CubicPath(rec["coeffs"].double()), func.double(), rec["z0"].double(), rec["t"].double(),
                         adjoint=False, method="dopri5", rtol=1e-10, atol=1e-12)
        func = func.float().cuda()
        func.nfe = 0
        X = tc.NaturalCubicSpline(rec["coeffs"].cuda())
        out = tc.cdeint(X, func, rec["z0"].cuda(), rec["t"].cuda(), adjoint=kw["adjoint"], method="dopri5",
                        rtol=kw["rtol"], atol=kw["atol"], options=dict(kw["options"]))
    assert out.shape == rec["out"].shape
    assert abs(func.nfe - rec["nfe"]) <= 0.35 * rec["nfe"]
    e_ref, e_gpu = rel(rec["out"], truth), rel(out, truth)
    assert e_gpu <= 3 * e_ref + 1e-6, (e_gpu, e_ref)


def test_cfg3_shape_against_oracle():
    """SpeechCommands-shaped: 161 steps, 21 channels, natural cubic, hidden 64, min_step 0.5, atol 1e-5, rtol 1e-3
    (src/ncde/ncde.py:129-134)."""
    import torchcde_b200 as tc
    from torchcde_b200 import adaptive
    torch.manual_seed(0)
    B, L, C, H = 64, 161, 21, 64
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, L, C, generator=g)
    x[..., 0] = torch.arange(L, dtype=torch.float32)
    x[..., 1:] = x[..., 1:].cumsum(-2) * 0.1
    func = O.SharedMLPField(C, H, H, 3)
    z0 = torch.randn(B, H, generator=g) * 0.5
    cref = O.natural_cubic_coeffs(x)
    Xr = O.CubicPath(cref)
    stats = {}
    with torch.no_grad():
        oref = O.cdeint(Xr, func, z0, Xr.interval, adjoint=False, method="dopri5", rtol=1e-3, atol=1e-5,
                        options={"min_step": 0.5}, stats=stats)
        c = tc.natural_cubic_coeffs(x.cuda())
        assert torch.equal(c.cpu(), cref)
        X = tc.NaturalCubicSpline(c)
        out = tc.cdeint(X, func.cuda(), z0.cuda(), X.interval, adjoint=False, method="dopri5", rtol=1e-3, atol=1e-5,
                        options={"min_step": 0.5})
        outb = tc.cdeint(X, func, z0.cuda(), X.interval, adjoint=False, method="dopri5", rtol=1e-3, atol=1e-5,
                         options={"min_step": 0.5, "precision": "bf16"})
    st = adaptive.last_stats
    assert abs(st["nfe"] - stats["nfe"]) <= 0.35 * stats["nfe"], (st, stats)
    assert rel(out, oref) <= 2e-2
    assert rel(outb, oref) <= 5e-2        # tensor-core tiles under the same controller


def test_dopri5_options_and_errors(problem):
    import torchcde_b200 as tc
    Xr, func, X, fd, z0 = problem
    t = torch.tensor([0., 2.5, 7.], dtype=torch.float64).cuda()
    with torch.no_grad():
        a = tc.cdeint(X, fd, z0.cuda(), t, adjoint=False, method="dopri5", rtol=1e-6, atol=1e-8)
        b = tc.cdeint(X, fd, z0.cuda(), t, adjoint=False, rtol=1e-6, atol=1e-8)            # method=None -> dopri5
        with pytest.raises(AssertionError, match="max_num_steps"):
            tc.cdeint(X, fd, z0.cuda(), t, adjoint=False, method="dopri5", rtol=1e-9, atol=1e-11,
                      options={"max_num_steps": 1})
        with pytest.warns(UserWarning):
            tc.cdeint(X, fd, z0.cuda(), t, adjoint=False, method="dopri5", options={"step_size": 1})
    assert torch.equal(a, b) and a.shape == (3, 3, 4)
    # backprop *through* the adaptive solver is not implemented (the continuous adjoint is: tests/test_gpu_adjoint.py)
    with pytest.raises(NotImplementedError):
        tc.cdeint(X, fd, z0.cuda().requires_grad_(True), t, adjoint=False, method="dopri5")
