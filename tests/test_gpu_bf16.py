"""The tensor-core path (options={'precision': 'bf16'}: bf16 operand tiles, fp32 accumulate, fp32 state) against
the oracle / the fp32 product path.  Stated bound (BASELINE.json north_star allows "a stated looser bound for
TF32/bf16 MLP tiles"), relative max-norm:
    short sequences (<= 11 knots):  hidden states 3e-3,  gradients 8e-2
    cfg-5 length (143 knots, 568 chained stages):  hidden states 1e-2,  gradients 1.5e-1
Every GEMM of the vector-field MLP (hidden layers and the final layer, forward and backward) takes bf16 operands —
8 mantissa bits, 2^-9 relative rounding — so the pre-activation of every vector-field evaluation carries ~1e-2 absolute
error; accumulation, the RK state, the path derivative and the parameter-gradient accumulators stay fp32."""
import copy

import pytest
import torch

from oracle import cde_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


# (1100, ...) and (700, ...): several 128-row tiles per CTA, so the double-buffered tile pipeline wraps around
CASES = [(1100, 3, 100, 128, 128, 3), (700, 3, 30, 16, 64, 2), (130, 4, 100, 128, 128, 3), (300, 6, 100, 128, 128, 3), (64, 5, 4, 64, 64, 3), (200, 5, 21, 64, 64, 2),
         (96, 4, 14, 32, 128, 1), (33, 3, 2, 32, 128, 0),
         # 128 channels (unpadded dX/dt pitch, one hidden row per group); 4-layer field (3 hidden GEMMs chained in one kernel)
         (200, 4, 128, 32, 64, 2), (150, 4, 10, 24, 48, 4)]


@pytest.mark.parametrize("B,L,C,H,HH,n", CASES)
def test_bf16_short_sequences_against_oracle(B, L, C, H, HH, n):
    import torchcde_b200 as tc
    g = torch.Generator().manual_seed(B + L + C)
    x = torch.randn(B, L, C, generator=g)
    x[..., 0] = torch.arange(L, dtype=torch.float32)
    x[..., 1:] = x[..., 1:].cumsum(-2) * 0.2
    torch.manual_seed(5)
    func = O.SharedMLPField(C, H, HH, n) if n > 0 else O.ToyField(C, H, width=HH)
    z0 = torch.randn(B, H, generator=g) * 0.5
    cref = O.linear_interpolation_coeffs(x.clone(), rectilinear=0)
    Xr = O.LinearPath(cref)
    w = torch.randn(B, cref.shape[1], H, generator=g)
    z0r = z0.clone().requires_grad_(True)
    oref = O.cdeint(Xr, func, z0r, Xr.grid_points, adjoint=False, method="rk4", options={"step_size": 1})
    (oref * w).sum().backward()
    gref = {k: p.grad.clone() for k, p in func.named_parameters()}
    fd = copy.deepcopy(func).cuda()
    for p in fd.parameters():
        p.grad = None
    X = tc.LinearInterpolation(cref.cuda())
    z0d = z0.cuda().requires_grad_(True)
    out = tc.cdeint(X, fd, z0d, X.grid_points, adjoint=False, method="rk4",
                    options={"step_size": 1, "precision": "bf16"})
    (out * w.cuda()).sum().backward()
    assert torch.isfinite(out).all()
    assert rel(out, oref) <= 3e-3
    assert rel(z0d.grad, z0r.grad) <= 8e-2
    for k, p in fd.named_parameters():
        assert torch.isfinite(p.grad).all(), k
        assert rel(p.grad, gref[k]) <= 8e-2, k


def test_bf16_euler_against_oracle():
    """Euler on the all-tensor-core path (one stage per step: the next stage input always comes from `advance`)."""
    import torchcde_b200 as tc
    g = torch.Generator().manual_seed(11)
    B, L, C, H, HH, n = 300, 6, 21, 64, 64, 2
    x = torch.randn(B, L, C, generator=g)
    x[..., 0] = torch.arange(L, dtype=torch.float32)
    x[..., 1:] = x[..., 1:].cumsum(-2) * 0.2
    torch.manual_seed(5)
    func = O.SharedMLPField(C, H, HH, n)
    z0 = torch.randn(B, H, generator=g) * 0.5
    cref = O.natural_cubic_coeffs(x)
    Xr = O.CubicPath(cref)
    w = torch.randn(B, L, H, generator=g)
    z0r = z0.clone().requires_grad_(True)
    oref = O.cdeint(Xr, func, z0r, Xr.grid_points, adjoint=False, method="euler", options={"step_size": 0.5})
    (oref * w).sum().backward()
    gref = {k: p.grad.clone() for k, p in func.named_parameters()}
    fd = copy.deepcopy(func).cuda()
    for p in fd.parameters():
        p.grad = None
    X = tc.NaturalCubicSpline(cref.cuda())
    z0d = z0.cuda().requires_grad_(True)
    out = tc.cdeint(X, fd, z0d, X.grid_points, adjoint=False, method="euler", options={"step_size": 0.5, "precision": "bf16"})
    (out * w.cuda()).sum().backward()
    assert rel(out, oref) <= 3e-3
    assert rel(z0d.grad, z0r.grad) <= 8e-2
    for k, p in fd.named_parameters():
        assert rel(p.grad, gref[k]) <= 8e-2, k


def test_bf16_full_length_against_fp32_path():
    import bench
    import torchcde_b200 as tc
    cfg = bench.CFG
    B = 256
    x, _, _ = bench.synth_batch(B, 7)
    coeffs = tc.linear_interpolation_coeffs(x.cuda(), rectilinear=0)
    torch.manual_seed(3)
    func = O.SharedMLPField(cfg["C"], cfg["H"], cfg["HH"], cfg["n_layers"])
    g = torch.Generator().manual_seed(4)
    z0 = torch.randn(B, cfg["H"], generator=g) * 0.5
    w = torch.randn(B, 143, cfg["H"], generator=g)
    res = {}
    for prec in ("fp32", "bf16"):
        fd = copy.deepcopy(func).cuda()
        X = tc.LinearInterpolation(coeffs)
        z = z0.cuda().requires_grad_(True)
        out = tc.cdeint(X, fd, z, X.grid_points, adjoint=False, method="rk4",
                        options={"step_size": 1, "precision": prec})
        (out * w.cuda()).sum().backward()
        res[prec] = (out.detach(), z.grad, {n: p.grad for n, p in fd.named_parameters()})
    o32, g32, p32 = res["fp32"]
    o16, g16, p16 = res["bf16"]
    assert rel(o16, o32) <= 1e-2
    assert rel(g16, g32) <= 1.5e-1
    for n in p32:
        assert rel(p16[n], p32[n]) <= 1.5e-1, n


def test_bf16_sub_batch_rows_are_bit_identical():
    """Rows are independent in the tensor-core path too: solving a 128-aligned sub-batch reproduces the rows exactly."""
    import torchcde_b200 as tc
    torch.manual_seed(0)
    B, L, C, H = 512, 10, 12, 64
    x = torch.randn(B, L, C)
    x[..., 0] = torch.arange(L, dtype=torch.float32)
    c = tc.linear_interpolation_coeffs(x.cuda())
    func = O.SharedMLPField(C, H, H, 2).cuda()
    z0 = (torch.randn(B, H) * 0.5).cuda()
    opts = {"step_size": 1, "precision": "bf16"}
    with torch.no_grad():
        X = tc.LinearInterpolation(c)
        full = tc.cdeint(X, func, z0, X.grid_points, adjoint=False, method="rk4", options=opts)
        Xs = tc.LinearInterpolation(c[128:384].contiguous())
        part = tc.cdeint(Xs, func, z0[128:384].contiguous(), Xs.grid_points, adjoint=False, method="rk4", options=opts)
    assert torch.equal(part, full[128:384])
