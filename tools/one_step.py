"""Runs warm-up + ONE bench step (cfg-5 shape) — the command profiled by ncu for profiles/*launches*.csv."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "online-neural-cdes_b200")]
import torch, bench, torchcde_b200 as tc, ncde_b200
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = bench.CFG
dev = torch.device("cuda")
torch.manual_seed(1)
model = ncde_b200.NeuralCDE(cfg["C"], cfg["H"], cfg["out"], static_dim=cfg["S"], hidden_hidden_dim=cfg["HH"],
                            num_layers=cfg["n_layers"], interpolation="rectilinear", adjoint=False, solver="rk4",
                            return_sequences=True, precision=prec).to(dev)
x, static, labels = bench.synth_batch(cfg["B"], 100)
coeffs = tc.linear_interpolation_coeffs(x.to(dev), rectilinear=0)
static, labels = static.to(dev), labels.to(dev)
lossf = torch.nn.BCEWithLogitsLoss()
for _ in range(steps):
    model.zero_grad()
    loss = lossf(model((static, coeffs)).squeeze(-1), labels)
    loss.backward()
torch.cuda.synchronize()
print("loss", float(loss.detach()))
