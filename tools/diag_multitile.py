"""Diagnostic: how much do the parameter gradients of a cfg-5 shard depend on the grouping of its tiles (1 x 1024, 2 x 512, 4 x 256, 8 x 128)?"""
import os, sys, copy
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "online-neural-cdes_b200"), os.path.join(R, "tests")]
import torch, bench, torchcde_b200 as tc
from oracle import cde_oracle as O
import parity_util as PU
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
cfg = bench.CFG
B = 1024
x, _, _ = bench.synth_batch(B, 11)
c = tc.linear_interpolation_coeffs(x.cuda(), rectilinear=0)
torch.manual_seed(5)
func = O.SharedMLPField(cfg["C"], cfg["H"], cfg["HH"], cfg["n_layers"])
g = torch.Generator().manual_seed(6)
z0 = (torch.randn(B, cfg["H"], generator=g) * 0.5).cuda()
w = torch.randn(B, c.shape[1], cfg["H"], generator=g).cuda()
def solve(rows, p=prec):
    fd = copy.deepcopy(func).cuda()
    X = tc.LinearInterpolation(c[rows].contiguous())
    z = z0[rows].clone().requires_grad_(True)
    out = tc.cdeint(X, fd, z, X.grid_points, adjoint=False, method="rk4", options={"step_size": 1, "precision": p})
    (out * w[rows]).sum().backward()
    return {n: q.grad.detach().double() for n, q in fd.named_parameters()}
res = {}
for T in (1024, 512, 256, 128):
    acc = None
    for k in range(B // T):
        gp = solve(slice(k * T, (k + 1) * T))
        acc = gp if acc is None else {n: acc[n] + gp[n] for n in gp}
    res[T] = acc
ref = solve(slice(0, B), "fp32")
for a, b in ((1024, 512), (512, 256), (256, 128), (1024, 128)):
    print(prec, "groups of %d vs %d:" % (a, b), {n.split(".")[0][:4] + n[-8:]: "%.1e" % PU.rel(res[a][n], res[b][n]) for n in res[a]})
for T in (1024, 128):
    print(prec, "groups of %d vs the fp32 path:" % T, {n.split(".")[0][:4] + n[-8:]: "%.1e" % PU.rel(res[T][n], ref[n]) for n in ref})
