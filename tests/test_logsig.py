"""Log-ODE transform, depth 1 to 3.  The reference's log-signature comes from the un-vendored `signatory` extension,
so no output of the reference itself can be minted here; the oracle's restatement (tensor-algebra log of the Chen signature,
coefficients at the Lyndon words = Signatory's default "words" mode) is pinned by three independent routes instead:
  * the algebraic identity of the reference's own test (modules/torchcde/test/test_log_ode.py:6-30): the derivative of the linearly
    interpolated transformed path at a window mid-point equals the log-signature of that window;
  * brute-force iterated sums over ordered segment tuples (no Chen recursion, no tensor algebra);
  * the Baker-Campbell-Hausdorff series for a two-segment path, log(e^a e^b) = a + b + [a,b]/2 + ([a,[a,b]] + [b,[b,a]])/12,
    evaluated with explicit commutators of tensors — a published closed form that shares no code with either of the above."""
import pytest
import torch

from oracle import cde_oracle as O


def _brute_signature(path, depth):
    """Iterated integrals of a piecewise-linear path as explicit sums over ordered tuples of segments:
    S2_ij = sum_{k<l} D_ki D_lj + 1/2 sum_k D_ki D_kj;  S3_ijk = sum_{a<b<c} + 1/2 (a=b<c) + 1/2 (a<b=c) + 1/6 (a=b=c)."""
    inc = (path[1:] - path[:-1]).double()
    m, d = inc.shape
    S1 = inc.sum(0)
    S2 = torch.zeros(d, d, dtype=torch.float64)
    S3 = torch.zeros(d, d, d, dtype=torch.float64)
    for a in range(m):
        for b in range(a, m):
            wab = 0.5 if a == b else 1.0
            S2 += wab * torch.outer(inc[a], inc[b])
            if depth >= 3:
                for c in range(b, m):
                    if a == b == c:
                        wt = 1.0 / 6.0
                    elif a == b or b == c:
                        wt = 0.5
                    else:
                        wt = 1.0
                    S3 += wt * torch.einsum("i,j,k->ijk", inc[a], inc[b], inc[c])
    return S1, S2, S3


def _brute_logsig3(path):
    S1, S2, S3 = _brute_signature(path, 3)
    d = path.size(-1)
    L2 = S2 - 0.5 * torch.outer(S1, S1)
    L3 = S3 - 0.5 * (torch.einsum("i,jk->ijk", S1, S2) + torch.einsum("ij,k->ijk", S2, S1)) + torch.einsum("i,j,k->ijk", S1, S1, S1) / 3.0
    out = [S1] + [L2[i, j].reshape(1) for (i, j) in O.lyndon_words(d, 2)] + [L3[i, j, k].reshape(1) for (i, j, k) in O.lyndon_words(d, 3)]
    return torch.cat(out).to(path.dtype)


def _brute_logsig(path, depth):
    """(m+1, d) -> depth-<=2 log-signature by explicit loops: A_ij = 1/2 sum_{k<l} (D_k,i D_l,j - D_k,j D_l,i)."""
    if depth == 3:
        return _brute_logsig3(path)
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


def test_lyndon_word_counts_and_order():
    for d in (1, 2, 3, 5, 14):
        assert len(O.lyndon_words(d, 1)) == d
        assert len(O.lyndon_words(d, 2)) == d * (d - 1) // 2
        assert len(O.lyndon_words(d, 3)) == (d ** 3 - d) // 3
        assert O.logsignature_channels(d, 3) == d + d * (d - 1) // 2 + (d ** 3 - d) // 3
    assert O.lyndon_words(3, 3) == [(0, 0, 1), (0, 0, 2), (0, 1, 1), (0, 1, 2), (0, 2, 1), (0, 2, 2), (1, 1, 2), (1, 2, 2)]


def test_depth3_matches_baker_campbell_hausdorff():
    """Two-segment path a then b: log S = a + b + [a,b]/2 + ([a,[a,b]] + [b,[b,a]])/12 with [x,y] = x(x)y - y(x)x."""
    torch.manual_seed(0)
    for d in (2, 3, 5):
        a, b = torch.randn(d, dtype=torch.float64), torch.randn(d, dtype=torch.float64)
        path = torch.stack([torch.zeros(d, dtype=torch.float64), a, a + b])
        ab = torch.outer(a, b) - torch.outer(b, a)                                     # [a, b], level 2
        a_ab = torch.einsum("i,jk->ijk", a, ab) - torch.einsum("ij,k->ijk", ab, a)      # [a, [a, b]]
        b_ba = torch.einsum("i,jk->ijk", b, -ab) - torch.einsum("ij,k->ijk", -ab, b)    # [b, [b, a]]
        L2, L3 = 0.5 * ab, (a_ab + b_ba) / 12.0
        expect = torch.cat([a + b] + [L2[i, j].reshape(1) for (i, j) in O.lyndon_words(d, 2)] +
                           [L3[i, j, k].reshape(1) for (i, j, k) in O.lyndon_words(d, 3)])
        got = O.logsignature_words(path, 3)
        assert got.shape == expect.shape and got.allclose(expect, atol=1e-13)


def test_depth3_matches_brute_force_and_depth2_closed_form():
    torch.manual_seed(1)
    for m, d in [(1, 3), (4, 2), (6, 4), (3, 1)]:
        path = torch.randn(m + 1, d, dtype=torch.float64).cumsum(0)
        got = O.logsignature_words(path, 3)
        assert got.allclose(_brute_logsig3(path), atol=1e-12)
        assert O.logsignature_words(path, 2).allclose(O.logsignature_depth2(path, 2), atol=1e-13)
    # a straight line has no area and no level-3 term
    line = torch.linspace(0, 1, 5, dtype=torch.float64).unsqueeze(-1) * torch.randn(3, dtype=torch.float64)
    assert O.logsignature_words(line, 3)[3:].abs().max() < 1e-14


@pytest.mark.parametrize("depth", [1, 2, 3])
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
                                 ((4, 16, 6), 1.7, 2, torch.linspace(0, 9, 16)), ((1, 9, 1), 3, 2, None),
                                 ((6, 24, 14), 4, 3, None), ((2, 2, 20, 4), 2.5, 3, None), ((3, 12, 1), 3, 3, None),
                                 ((4, 16, 5), 1.7, 3, torch.linspace(0, 9, 16))]:
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
        tc.logsig_windows(torch.randn(2, 8, 3).cuda(), 4, 2)
    assert tc.logsignature_channels(14, 2) == 105 and tc.logsignature_channels(14, 3) == 105 + 910


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
