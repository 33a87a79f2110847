"""Log-ODE transform.  The reference's log-signature comes from the un-vendored `signatory` extension, so the
oracle's restatement is "parity unpinned"; what pins it is the algebraic identity of the reference's own test
(modules/torchcde/test/test_log_ode.py:6-30): the derivative of the linearly interpolated transformed path at a window
mid-point equals the log-signature of that window — checked here against an independent brute-force Levy area."""
import pytest
import torch

from oracle import cde_oracle as O


def _brute_logsig(path, depth):
    """(m+1, d) -> depth-<=2 log-signature by explicit loops: A_ij = 1/2 sum_{k<l} (D_k,i D_l,j - D_k,j D_l,i)."""
    inc = path[1:] - path[:-1]
    d = path.size(-1)
    out = [float(inc[:, i].sum()) for i in range(d)]
    if depth == 2:
        m = inc.size(0)
        for i in range(d):
            for j in range(i + 1, d):
                a = 0.0
                for k in range(m):
                    for l in range(k + 1, m):
                        a += float(inc[k, i] * inc[l, j] - inc[k, j] * inc[l, i])
                out.append(0.5 * a)
    return torch.tensor(out, dtype=path.dtype)


@pytest.mark.parametrize("depth", [1, 2])
def test_oracle_identity_of_the_reference_test(depth):
    torch.manual_seed(depth)
    window_length = 4
    for pieces in (1, 2, 3, 5, 10):
        num_channels = int(torch.randint(low=1, high=4, size=(1,)))
        x_ = [torch.randn(1, num_channels, dtype=torch.float64)]
        expect = []
        for _ in range(pieces):
            x = torch.randn(window_length, num_channels, dtype=torch.float64)
            expect.append(_brute_logsig(torch.cat([x_[-1][-1:], x]), depth))
            x_.append(x)
        x = torch.cat(x_)
        logsig_x = O.logsig_windows(x, depth, window_length)
        assert logsig_x.shape == (pieces + 1, O.logsignature_channels(num_channels, depth))
        X = O.LinearPath(O.linear_interpolation_coeffs(logsig_x))
        point = 0.5
        for e in expect:
            assert X.derivative(torch.tensor(point, dtype=torch.float64)).allclose(e)
            point += 1


def test_oracle_windows_with_missing_values_and_fractional_windows():
    torch.manual_seed(3)
    x = torch.randn(3, 2, 11, 4, dtype=torch.float64)
    x[0, 0, 3, 1] = float("nan")
    x[1, 1, 7, :] = float("nan")
    t = torch.linspace(0, 5, 11, dtype=torch.float64)
    out = O.logsig_windows(x, 2, 1.3, t)
    assert out.shape == (3, 2, 5, 10) and torch.isfinite(out).all()   # ceil(5 / 1.3) = 4 windows
    # the level-1 part telescopes: last row = x_end
    filled = O.linear_interpolation_coeffs(x, t)
    assert out[..., -1, :4].allclose(filled[..., -1, :])
    vals, times = O.logsig_windows(x, 2, 1.3, t, _version=0)
    assert times.shape == (5,) and float(times[-1]) == 5.0


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 2e-5)])
def test_gpu_logsig_windows_matches_oracle(dtype, tol):
    import torchcde_b200 as tc
    torch.manual_seed(5)
    for shape, wl, depth, tt in [((7, 24, 14), 4, 2, None), ((2, 3, 25, 5), 2.5, 2, None), ((5, 30, 3), 7, 1, None),
                                 ((4, 16, 6), 1.7, 2, torch.linspace(0, 9, 16)), ((1, 9, 1), 3, 2, None)]:
        x = torch.randn(*shape, dtype=dtype)
        x[..., 0] = torch.arange(shape[-2], dtype=dtype) if shape[-1] > 1 else x[..., 0]
        if shape[-1] > 2:
            x[..., 5, 2] = float("nan")
        t = None if tt is None else tt.to(dtype)
        ref = O.logsig_windows(x.clone(), depth, wl, t)
        got = tc.logsig_windows(x.clone().cuda(), depth, wl, None if t is None else t.cuda())
        assert got.shape == ref.shape
        assert float((got.cpu() - ref).abs().max() / ref.abs().max().clamp_min(1e-30)) <= tol
        rv, rt = O.logsig_windows(x.clone(), depth, wl, t, _version=0)
        gv, gt = tc.logsignature_windows(x.clone().cuda(), depth, wl, None if t is None else t.cuda())
        assert torch.equal(gt.cpu(), rt) and float((gv.cpu() - rv).abs().max() / rv.abs().max().clamp_min(1e-30)) <= tol
    with pytest.raises(NotImplementedError):
        tc.logsig_windows(torch.randn(2, 8, 3).cuda(), 3, 2)
    assert tc.logsignature_channels(14, 2) == 105


@pytest.mark.gpu
def test_gpu_cfg4_logsig_path_feeds_the_solver():
    """cfg 4 shape: 13 channels + time, 24 steps, depth-2 log-signature (105 channels), online outputs."""
    import copy
    import torchcde_b200 as tc
    torch.manual_seed(8)
    B, L, d, H = 32, 24, 14, 32
    x = torch.randn(B, L, d) * 0.3
    x[..., 0] = torch.arange(L, dtype=torch.float32)
    ls_ref = O.logsig_windows(x.clone(), 2, 4)
    ls = tc.logsig_windows(x.clone().cuda(), 2, 4)
    assert float((ls.cpu() - ls_ref).abs().max()) <= 1e-4
    func = O.SharedMLPField(105, H, H, 2)
    z0 = torch.randn(B, H) * 0.5
    Xr = O.LinearPath(O.linear_interpolation_coeffs(ls_ref))
    with torch.no_grad():
        ref = O.cdeint(Xr, func, z0, Xr.grid_points, adjoint=False, method="rk4", options={"step_size": 1})
    X = tc.LinearInterpolation(tc.linear_interpolation_coeffs(ls))
    with torch.no_grad():
        out = tc.cdeint(X, copy.deepcopy(func).cuda(), z0.cuda(), X.grid_points, adjoint=False, method="rk4",
                        options={"step_size": 1, "precision": "fp32"})
    assert float((out.cpu() - ref).abs().max() / ref.abs().max()) <= 1e-4
