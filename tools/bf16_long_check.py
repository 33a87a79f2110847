"""bf16 tensor-core path vs fp32 product path on the full cfg-5 length (143 knots): error growth over 568 stages."""
import os, sys, copy
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "online-neural-cdes_b200")]
import torch, bench
import torchcde_b200 as tc
from oracle import cde_oracle as O
cfg = bench.CFG
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
x, static, labels = bench.synth_batch(B, 7)
dev = torch.device("cuda")
coeffs = tc.linear_interpolation_coeffs(x.to(dev), rectilinear=0)
torch.manual_seed(3)
func = O.SharedMLPField(cfg["C"], cfg["H"], cfg["HH"], cfg["n_layers"])
g = torch.Generator().manual_seed(4)
z0 = torch.randn(B, cfg["H"], generator=g) * 0.5
w = torch.randn(B, 143, cfg["H"], generator=g)
res = {}
for prec in ("fp32", "bf16"):
    fd = copy.deepcopy(func).to(dev)
    X = tc.LinearInterpolation(coeffs)
    z = z0.to(dev).requires_grad_(True)
    out = tc.cdeint(X, fd, z, X.grid_points, adjoint=False, method="rk4", options={"step_size": 1, "precision": prec})
    (out * w.to(dev)).sum().backward()
    res[prec] = (out.detach(), z.grad, {n: p.grad for n, p in fd.named_parameters()})
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
o32, g32, p32 = res["fp32"]; o16, g16, p16 = res["bf16"]
print("out", rel(o16, o32), "out@last", rel(o16[:, -1], o32[:, -1]), "z0", rel(g16, g32), {n: "%.1e" % rel(p16[n], p32[n]) for n in p32})
print("out rms rel", float((o16 - o32).pow(2).mean().sqrt() / o32.pow(2).mean().sqrt()), "grad z0 rms rel", float((g16 - g32).pow(2).mean().sqrt() / g32.pow(2).mean().sqrt()))
