"""Host enqueue time vs GPU time of one cfg-5 step (is the launch rate the limiter?)."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "online-neural-cdes_b200")]
import torch, bench, torchcde_b200 as tc, ncde_b200
cfg = bench.CFG
dev = torch.device("cuda")
torch.manual_seed(1)
model = ncde_b200.NeuralCDE(cfg["C"], cfg["H"], cfg["out"], static_dim=cfg["S"], hidden_hidden_dim=cfg["HH"],
                            num_layers=cfg["n_layers"], interpolation="rectilinear", adjoint=False, solver="rk4",
                            return_sequences=True, precision="bf16").to(dev)
x, static, labels = bench.synth_batch(cfg["B"], 100)
coeffs = tc.linear_interpolation_coeffs(x.to(dev), rectilinear=0)
static, labels = static.to(dev), labels.to(dev)
lossf = torch.nn.BCEWithLogitsLoss()
for it in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    model.zero_grad()
    out = model((static, coeffs))
    t1 = time.perf_counter()
    loss = lossf(out.squeeze(-1), labels)
    loss.backward()
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    if it >= 2:
        print("host enqueue fwd %.1f ms, bwd %.1f ms, total wall %.1f ms (GPU drained %.1f ms after the last enqueue)" %
              ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t0) * 1e3, (t3 - t2) * 1e3))
