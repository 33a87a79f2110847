"""Throughput of the widened vector-field modes on the fp32 fixed-grid path, with the CPU oracle timed beside
them: fwd+bwd seq-steps/s for {original, minimal, gru} x {matmul, evaluate, derivative} and the cost of the path gradient.

    python tools/bench_modes.py [--batch 1024] [--steps 3] [--cpu-only]        -> one JSON line per mode

Workload: cfg-2-shaped (CharacterTrajectories-like: 182 knots, 4 channels incl. time, hidden 64, hidden-hidden 64, 3 layers),
rk4 step 1, backprop through the solver (`adjoint=False`, what the reference's `sparsity` ablation runs).
Algorithmic FLOPs per seq-step: 12 * F_eval with F_eval of the mode's own layer shapes.
"""
import argparse, json, os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "online-neural-cdes_b200")]
import torch
from oracle import cde_oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--cpu-batch", type=int, default=64)
ap.add_argument("--cpu-only", action="store_true")
args = ap.parse_args()
K, C, H, HH, n = 182, 4, 64, 64, 3
FIELDS = {"original": O.SharedMLPField, "minimal": O.MinimalGatedField, "gru": O.GRUGatedField}


def f_eval(kind, vft):
    d0 = H if vft == "matmul" else H + C
    out = H * C if vft == "matmul" else H
    hidden = d0 * HH + (n - 1) * HH * HH
    if kind == "gru":
        hidden = 2 * hidden + d0 * d0
    heads = (2 if kind != "original" else 1) * HH * out
    return 2.0 * (hidden + heads) + (2.0 * H * C if vft == "matmul" else 0.0)


def make(kind, vft, B, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, K, C, generator=g).cumsum(-2) * 0.1
    x[..., 0] = torch.arange(K, dtype=torch.float32)
    torch.manual_seed(seed)
    func = FIELDS[kind](C, H, HH, n, vector_field_type=vft)
    z0 = torch.randn(B, H, generator=g) * 0.5
    return x, func, z0


def cpu_rate(kind, vft):
    x, func, z0 = make(kind, vft, args.cpu_batch)
    X = O.LinearPath(x)
    t0 = time.perf_counter()
    z0r = z0.clone().requires_grad_(True)
    out = O.cdeint(X, func, z0r, X.interval, adjoint=False, method="rk4", options={"step_size": 1}, vector_field_type=vft)
    out[:, -1].sum().backward()
    dt = time.perf_counter() - t0
    return args.cpu_batch * (K - 1) / dt, dt


for kind in FIELDS:
    for vft in ("matmul", "evaluate", "derivative"):
        cpu_v, cpu_s = cpu_rate(kind, vft)
        line = {"metric": "ncde_fwd_bwd_seq_steps_per_sec", "unit": "seq-steps/s", "vector_field": kind, "vector_field_type": vft,
                "config": {"workload": "cfg2_shaped_fp32_modes", "batch": args.batch, "knots": K, "channels": C, "hidden": H,
                           "hidden_hidden": HH, "layers": n, "solver": "rk4 step 1, backprop through the solver", "precision": "fp32"},
                "cpu_baseline": {"value": cpu_v, "unit": "seq-steps/s", "cores": torch.get_num_threads(), "kind": "port",
                                 "sample": "oracle, %d series, all %d steps, fwd+bwd, %.1f s" % (args.cpu_batch, K - 1, cpu_s)}}
        if not args.cpu_only:
            import torchcde_b200 as tc
            x, func, z0 = make(kind, vft, args.batch)
            func = func.cuda()
            coeffs = x.cuda()
            z0d = z0.cuda().requires_grad_(True)
            X = tc.LinearInterpolation(coeffs)

            def step(path_grad=False):
                Xs = X
                if path_grad:
                    Xs = tc.LinearInterpolation(coeffs.detach().clone().requires_grad_(True))
                out = tc.cdeint(Xs, func, z0d, Xs.interval, adjoint=False, vector_field_type=vft, method="rk4",
                                options={"step_size": 1})
                out[:, -1].sum().backward()

            def timed(path_grad=False):
                for _ in range(2):
                    step(path_grad)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); a.record()
                for _ in range(args.steps):
                    step(path_grad)
                b.record(); torch.cuda.synchronize()
                return a.elapsed_time(b) / args.steps

            ms = timed()
            line.update(value=args.batch * (K - 1) / (ms * 1e-3), ms_per_step=ms,
                        algorithmic_tflops=12.0 * f_eval(kind, vft) * args.batch * (K - 1) / (ms * 1e-3) / 1e12)
            if kind == "original" and vft == "matmul":
                ms_pg = timed(True)
                line["with_path_gradient"] = {"ms_per_step": ms_pg, "overhead": ms_pg / ms - 1.0}
        print(json.dumps(line), flush=True)
