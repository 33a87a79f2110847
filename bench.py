#!/usr/bin/env python
"""Benchmark of the Neural-CDE solve hot path (BASELINE.json metric: NCDE fwd+bwd sequence-steps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg5] [--precision bf16x3|bf16|fp32]
                    [--global-batch G] [--impl b200|reference]

Workloads (--config; BASELINE.json `configs`, SURVEY.md §8 sizes; synthetic data of the named shapes):
  cfg1        toy Brownian paths, 128 series, L=3 -> 5 rectilinear knots, 2 channels, hidden 32, width 128, RK4, online MSE
  cfg2_linear CharacterTrajectories-shaped: 1024 series, L=182, 3+time channels, hidden 64, linear, RK4, 20-class CE
  cfg2_rect   the same data, rectilinear (363 knots)
  cfg3        SpeechCommands-shaped: 256 series, L=161, 20+time channels, hidden 64, natural cubic, dopri5 + continuous
              adjoint (reports NFE and accepted / attempted steps), 35-class CE
  cfg4        Beijing-PM2.5-shaped: 1024 series, L=24, 13+time channels -> depth-2 log-signatures over windows of 2 (105
              channels, inside the timed step), hidden 64, RK4, online RMSE
  cfg5        (default; the configuration the metric is quoted on) MIMIC-IV-shaped online sepsis: per GPU 1024 series (8192
              over 8 GPUs), 72 hourly steps -> 143 rectilinear knots, 100 channels, 5 static features, hidden =
              hidden-hidden = 128, 3 vector-field layers, 3/8-rule RK4 step 1, online BCE, backprop through the solver, Adam

One "step" = one forward + backward + optimiser pass over one batch.  seq-steps = B * (K - 1) per step.
Prints ONE JSON line on rank 0.  `--impl reference` runs the UNMODIFIED reference (torchcde.cdeint of oracle/_ref, staged by
oracle/make_ref.sh) on the host cores on the same config and batch.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "online-neural-cdes_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

CONFIGS = {
    "cfg1": dict(name="cfg1_sim_bm_toy", B=128, L=3, C=2, S=0, H=32, HH=128, n_layers=0, out=1, field="toy",
                 interp="rectilinear", method="rk4", adjoint=False, online=True, loss="mse", obs_rate=1.0),
    "cfg2_linear": dict(name="cfg2_character_trajectories_shaped_linear", B=1024, L=182, C=4, S=0, H=64, HH=64, n_layers=3,
                        out=20, field="orig", interp="linear", method="rk4", adjoint=False, online=False, loss="ce",
                        obs_rate=1.0),
    "cfg2_rect": dict(name="cfg2_character_trajectories_shaped_rectilinear", B=1024, L=182, C=4, S=0, H=64, HH=64,
                      n_layers=3, out=20, field="orig", interp="rectilinear", method="rk4", adjoint=False, online=False,
                      loss="ce", obs_rate=0.7),
    "cfg3": dict(name="cfg3_speech_commands_shaped_cubic_dopri5_adjoint", B=256, L=161, C=21, S=0, H=64, HH=64, n_layers=3,
                 out=35, field="orig", interp="cubic", method="dopri5", adjoint=True, online=False, loss="ce", obs_rate=1.0),
    "cfg4": dict(name="cfg4_beijing_pm25_shaped_logsig_depth2", B=1024, L=24, C=14, S=0, H=64, HH=64, n_layers=3, out=1,
                 field="orig", interp="linear", method="rk4", adjoint=False, online=True, loss="rmse", obs_rate=0.9,
                 logsig=dict(depth=2, window=2)),
    "cfg5": dict(name="cfg5_mimic_iv_shaped_online_sepsis", B=1024, L=72, C=100, S=5, H=128, HH=128, n_layers=3, out=1,
                 field="orig", interp="rectilinear", method="rk4", adjoint=False, online=True, loss="bce", obs_rate=0.2),
}
CFG = CONFIGS["cfg5"]   # the configuration the metric is quoted on (tests import this)
METRIC = "ncde_fwd_bwd_seq_steps_per_sec"
UNIT = "seq-steps/s"
DTYPES = {"fp32": "f32", "bf16": "bf16 tiles, f32 accumulate/state",
          "bf16x3": "bf16 hi+lo split tiles (3 tensor-core MMAs per GEMM), f32 accumulate/state"}
# relative max-norm against the oracle at this configuration's full length (tests/test_gpu_cfg5_full.py states and checks them)
PARITY_BOUNDS = {"fp32": {"states": 1e-5, "gradients": 1e-5},
                 "bf16x3": {"states": 1e-4, "gradients_smooth_field": 2e-4, "gradients_relu_field_median_row": 1e-4,
                            "gradients_relu_field_max": 1.5e-2,
                            "note": "measured 1.0e-5 / 9e-5 / 7e-6 / 7e-3; the max over a ReLU field is set by ReLU branch flips "
                                    "(the reference's own fp32 vs fp64 gradient differs by 1.7e-4 there)"},
                 "bf16": {"states": 1e-2, "gradients": 1.5e-1}}


def synth_batch(B, seed, cfg=CFG):
    """Seeded synthetic batch of the config's shape: time channel = arange (get_data/common.py:178-184), z-normalised
    values observed with probability obs_rate (NaN elsewhere), first row NaN -> 0 (transformers.py:53-55)."""
    g = torch.Generator().manual_seed(seed)
    L, C = cfg["L"], cfg["C"]
    if cfg["field"] == "toy":   # Brownian increments N(0, dt) (sim_bm_toy_example.py:66-83)
        dt = 1.0 / (L - 1)
        bm = torch.cat([torch.zeros(B, 1, 1), (torch.randn(B, L - 1, 1, generator=g) * dt ** 0.5).cumsum(1)], 1)
        x = torch.cat([torch.linspace(0, 1, L).view(1, L, 1).expand(B, L, 1), bm], -1).contiguous()
        return x, torch.zeros(B, 0), bm[..., 0].contiguous()
    x = torch.randn(B, L, C, generator=g)
    x[..., 0] = torch.arange(L, dtype=torch.float32)
    if cfg["obs_rate"] < 1.0:
        miss = torch.rand(B, L, C, generator=g) > cfg["obs_rate"]
        miss[..., 0] = False
        x[miss] = float("nan")
        first = x[:, 0, :]
        first[torch.isnan(first)] = 0.0
    static = torch.randn(B, cfg["S"], generator=g)
    if cfg["loss"] == "bce":
        labels = (torch.rand(B, L, generator=g) < 0.1).float()
    elif cfg["loss"] == "ce":
        labels = torch.randint(0, cfg["out"], (B,), generator=g)
    else:
        labels = torch.randn(B, n_outputs(cfg), generator=g)
    return x, static, labels


def knots(cfg):
    if "logsig" in cfg:
        return -(-(cfg["L"] - 1) // cfg["logsig"]["window"]) + 1
    return 2 * cfg["L"] - 1 if cfg["interp"] == "rectilinear" else cfg["L"]


def n_outputs(cfg):
    if not cfg["online"]:
        return 1
    return cfg["L"] if cfg["interp"] == "rectilinear" else knots(cfg)


def path_channels(cfg):
    if "logsig" in cfg:
        d = cfg["C"]
        return d + d * (d - 1) // 2
    return cfg["C"]


def loss_fn(cfg, out, labels):
    if cfg["loss"] == "bce":
        return torch.nn.functional.binary_cross_entropy_with_logits(out.squeeze(-1), labels)
    if cfg["loss"] == "ce":
        return torch.nn.functional.cross_entropy(out, labels)
    if cfg["loss"] == "rmse":
        return ((out.squeeze(-1) - labels) ** 2).mean().sqrt()
    return ((out.squeeze(-1) - labels) ** 2).mean()


def config_dict(cfg, B, world, scaling):
    """Identical in both arms (the driver compares it)."""
    solver = "rk4(3/8) step 1, backprop through solver" if cfg["method"] == "rk4" else \
        "dopri5 rtol 1e-3 atol 1e-5 min_step 0.5, continuous adjoint"
    return {"workload": cfg["name"], "batch_per_gpu": B, "global_batch": B * world, "length": cfg["L"], "knots": knots(cfg),
            "channels": path_channels(cfg), "static": cfg["S"], "hidden": cfg["H"], "hidden_hidden": cfg["HH"],
            "vector_field_layers": cfg["n_layers"], "interpolation": cfg["interp"], "solver": solver + ", Adam",
            "online": cfg["online"], "scaling_mode": scaling,
            "l2": "256 MB flush between timed steps (GPU arm)"}


def field_flops_per_sample(cfg=CFG):
    """Algorithmic FLOPs of the final layer + contraction per vector-field evaluation per sample."""
    C = path_channels(cfg)
    return 2.0 * cfg["HH"] * cfg["H"] * C + 2.0 * cfg["H"] * C


def eval_flops_per_sample(cfg=CFG):
    H, HH, C, n = cfg["H"], cfg["HH"], path_channels(cfg), cfg["n_layers"]
    if cfg["field"] == "toy":
        return 2.0 * (H * H + H * HH + HH * H * C) + 2.0 * H * C
    return 2.0 * (H * HH + (n - 1) * HH * HH + HH * H * C) + 2.0 * H * C


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(tensor_tflops=d.get("bf16_tflops_sustained", d.get("bf16_tflops")), hbm_gbs=d.get("hbm_gbs"),
                    source="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(tensor_tflops=1400.0, hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active," \
        "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # median over the samples taken under load (an idle sample reads the idle clock)
        busy = sorted(v for v in sm if mx is None or v >= 0.6 * mx) or sorted(sm)
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the UNMODIFIED reference (oracle/_ref: torchcde + torchdiffeq + src/ncde vector field, staged by
# oracle/make_ref.sh) when present, else the oracle port — on the host cores
# ----------------------------------------------------------------------------------------------------------------
def _reference_modules():
    ref = os.path.join(ROOT, "oracle", "_ref")
    if os.path.isdir(os.path.join(ref, "torchcde")):
        if ref not in sys.path:
            sys.path.insert(0, ref)
        import torchcde  # the reference's own package
        from ncde_ref_vector_fields.base import OriginalVectorField
        return "reference", torchcde, OriginalVectorField
    return "port", None, None


class _ToyFieldCPU(torch.nn.Module):
    """experiments/sim_bm_toy_example.py:10-30 (CDEFunc), restated for the reference arm: the script itself needs matplotlib."""

    def __init__(self, C, H, width):
        super().__init__()
        self.C, self.H = C, H
        self.linear0, self.linear1 = torch.nn.Linear(H, H), torch.nn.Linear(H, width)
        self.linear2 = torch.nn.Linear(width, C * H)

    def forward(self, t, z):
        z = self.linear1(self.linear0(z).relu()).relu()
        return self.linear2(z).tanh().view(z.size(0), self.H, self.C)


def cpu_reference_step_fn(cfg, B, seed=0):
    from oracle import cde_oracle as O
    kind, rtc, RefField = _reference_modules()
    torch.manual_seed(1)
    C = path_channels(cfg)
    if cfg["field"] == "toy":
        func = _ToyFieldCPU(C, cfg["H"], cfg["HH"])
    elif kind == "reference":
        func = RefField(C, cfg["H"], cfg["HH"], cfg["n_layers"])
    else:
        func = O.SharedMLPField(C, cfg["H"], cfg["HH"], cfg["n_layers"])
    initial = torch.nn.Linear(C + cfg["S"], cfg["H"])
    readout = torch.nn.Linear(cfg["H"], cfg["out"])
    params = list(func.parameters()) + list(initial.parameters()) + list(readout.parameters())
    opt = torch.optim.Adam(params, lr=1e-3)
    x, static, labels = synth_batch(B, seed, cfg)
    lib = rtc if kind == "reference" else None
    note = "unmodified reference: torchcde.cdeint + torchdiffeq from oracle/_ref" if lib else \
        "oracle/cde_oracle.py (PyTorch CPU restatement of the reference path)"

    def coeffs_of(xx):
        if "logsig" in cfg:   # `signatory` is un-vendored: the windows come from the oracle's restatement in both cases
            xx = O.logsig_windows(xx, cfg["logsig"]["depth"], cfg["logsig"]["window"])
        if cfg["interp"] == "cubic":
            return (lib or O).natural_cubic_coeffs(xx)
        return (lib or O).linear_interpolation_coeffs(xx, rectilinear=0 if cfg["interp"] == "rectilinear" else None)

    coeffs = coeffs_of(x) if "logsig" not in cfg else None
    stats = {}

    def step():
        opt.zero_grad(set_to_none=True)
        c = coeffs if coeffs is not None else coeffs_of(x)
        if lib:
            X = lib.NaturalCubicSpline(c) if cfg["interp"] == "cubic" else lib.LinearInterpolation(c)
        else:
            X = O.CubicPath(c) if cfg["interp"] == "cubic" else O.LinearPath(c)
        x0 = X.evaluate(X.interval[0] if cfg["field"] == "toy" else 0)
        h0 = initial(torch.cat((static, x0), -1) if cfg["S"] else x0)
        t = X.grid_points if cfg["online"] else X.interval
        kw = dict(method=cfg["method"], adjoint=cfg["adjoint"], rtol=1e-3, atol=1e-5,
                  options={"step_size": 1} if cfg["method"] == "rk4" else {"min_step": 0.5})
        if hasattr(func, "nfe"):
            func.nfe = 0
        hidden = (lib or O).cdeint(X, func, h0, t, **kw)
        if cfg["online"]:
            out = readout(hidden)
            out = out[:, ::2] if cfg["interp"] == "rectilinear" else out
        else:
            out = readout(hidden[:, -1])
        loss = loss_fn(cfg, out, labels)
        loss.backward()
        opt.step()
        stats["nfe"] = getattr(func, "nfe", None)
        return float(loss.detach())

    return step, B * (knots(cfg) - 1), kind, note, stats


def time_cpu(cfg, B, budget_s=20.0):
    step, units, kind, note, _ = cpu_reference_step_fn(cfg, B)
    tiny = cpu_reference_step_fn(cfg, min(B, 4))[0]
    tiny()  # lazy initialisation outside the timing
    t, spent = [], 0.0
    while spent < budget_s and len(t) < 3:
        t0 = time.perf_counter()
        step()
        t.append(time.perf_counter() - t0)
        spent += t[-1]
    return units / min(t), min(t), len(t), kind, note


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    B = args.batch_per_gpu
    step, units, kind, note, stats = cpu_reference_step_fn(cfg, B)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = units / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(cfg, B, args.gpus, args.scaling),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": "%s; all %d series of one GPU's batch per step, all %d knot intervals, fwd+bwd+Adam"
                                       % (note, B, knots(cfg) - 1)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if stats.get("nfe") is not None:
        line["nfe_forward_per_step"] = stats["nfe"]
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
class _ToyFieldGPU(torch.nn.Module):
    """CDEFunc of experiments/sim_bm_toy_example.py:10-30 (lowered by torchcde_b200.lowering through torch.fx)."""

    def __init__(self, C, H, width):
        super().__init__()
        self.input_channels, self.hidden_channels = C, H
        self.linear0, self.linear1 = torch.nn.Linear(H, H), torch.nn.Linear(H, width)
        self.linear2 = torch.nn.Linear(width, C * H)

    def forward(self, t, z):
        z = self.linear0(z).relu()
        z = self.linear1(z).relu()
        z = self.linear2(z).tanh()
        return z.view(z.size(0), self.hidden_channels, self.input_channels)


class _ToyModel(torch.nn.Module):
    def __init__(self, cfg, precision):
        super().__init__()
        self.func = _ToyFieldGPU(cfg["C"], cfg["H"], cfg["HH"])
        self.initial = torch.nn.Linear(cfg["C"], cfg["H"])
        self.readout = torch.nn.Linear(cfg["H"], cfg["out"])
        self.precision = precision

    def forward(self, coeffs):
        import torchcde_b200 as tc
        X = tc.LinearInterpolation(coeffs)
        z0 = self.initial(X.evaluate(X.interval[0]))
        gp = X.grid_points
        z = tc.cdeint(X, self.func, z0, gp, adjoint=False, method="rk4",
                      options={"step_size": 1, "precision": self.precision})
        return self.readout(z[:, 0::2])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg5", choices=sorted(CONFIGS))
    ap.add_argument("--precision", default=os.environ.get("NCDE_PRECISION", "default"))
    ap.add_argument("--batch-per-gpu", type=int, default=None)
    ap.add_argument("--global-batch", type=int, default=None,
                    help="strong scaling: this many series in total, split evenly over the GPUs (cfg 5: 8192)")
    ap.add_argument("--cpu-baseline-seconds", type=float, default=20.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-alt-mode", action="store_true", help="skip timing the other tensor-core precision mode")
    ap.add_argument("--check", action="store_true",
                    help="multi-GPU correctness: all-reduced N-rank gradients == single-rank gradients of the concatenated batch")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    args.scaling = "strong" if args.global_batch else "weak"
    if args.global_batch:
        assert args.global_batch % max(args.gpus, 1) == 0
        args.batch_per_gpu = args.global_batch // max(args.gpus, 1)
    if args.batch_per_gpu is None:
        args.batch_per_gpu = cfg["B"]
    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return
    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"

    import torch.distributed as dist
    import torchcde_b200 as tc
    import ncde_b200
    from torchcde_b200 import _capi, solver
    from torchcde_b200.distributed import allreduce_gradients

    if args.precision == "default":
        args.precision = "bf16x3" if "bf16x3" in solver._PRECISIONS else "bf16"
    if cfg["method"] == "dopri5" and args.precision == "bf16x3":
        args.precision = "bf16"
    assert args.precision in solver._PRECISIONS, sorted(solver._PRECISIONS)

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.check:
        run_check(args, cfg, rank, world, dev)
        if world > 1:
            dist.destroy_process_group()
        return
    B = args.batch_per_gpu
    K = knots(cfg)

    # model: identical initial weights on every rank
    torch.manual_seed(1)
    if cfg["field"] == "toy":
        model = _ToyModel(cfg, args.precision).to(dev)
    else:
        model = ncde_b200.NeuralCDE(path_channels(cfg), cfg["H"], cfg["out"], static_dim=cfg["S"] or None,
                                    hidden_hidden_dim=cfg["HH"], num_layers=cfg["n_layers"], interpolation=cfg["interp"],
                                    adjoint=cfg["adjoint"], solver=cfg["method"], return_sequences=cfg["online"],
                                    precision=args.precision).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)

    # data: this rank's shard of the global batch.  Coefficients are built once on the device (offline in the reference); the
    # log-signature transform of cfg 4 is part of the timed step.
    x, static_h, labels_h = synth_batch(B, seed=100 + rank, cfg=cfg)

    def coeffs_of(xd):
        if "logsig" in cfg:
            xd = tc.logsig_windows(xd, cfg["logsig"]["depth"], cfg["logsig"]["window"])
        if cfg["interp"] == "cubic":
            return tc.natural_cubic_coeffs(xd)
        return tc.linear_interpolation_coeffs(xd, rectilinear=0 if cfg["interp"] == "rectilinear" else None)

    raw_in_step = "logsig" in cfg
    src = x.to(dev) if raw_in_step else coeffs_of(x.to(dev))
    static, labels = static_h.to(dev), labels_h.to(dev)
    src_h = src.cpu().pin_memory()
    static_h, labels_h = static_h.pin_memory(), labels_h.pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    launches = {"n": 0}
    from torchcde_b200 import adaptive

    def step(c, s, y):
        opt.zero_grad(set_to_none=True)
        if raw_in_step:
            c = coeffs_of(c)
        out = model((s, c)) if cfg["S"] else model(c)
        loss = loss_fn(cfg, out, y)
        loss.backward()
        if world > 1:
            allreduce_gradients(model.parameters(), average=True)
        opt.step()
        # kernels of libncde_b200 this step: solve fwd + bwd, spline ctor, X(0) (+ the log-signature / fill kernels of cfg 4)
        launches["n"] += solver.last_launches["fwd"] + solver.last_launches["bwd"] + 2 + (3 if raw_in_step else 0)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(src, static, labels)
    barrier()

    # ---- device-resident timing: K steps, CUDA events per step, L2 flushed between steps ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches["n"] = 0
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in ev:
        flush.zero_()
        a.record()
        step(src, static, labels)
        b.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_per_step = float(tt.item()) / args.steps
    gpu_launches = launches["n"]
    units_per_step = world * B * (K - 1)
    value = units_per_step / (ms_per_step * 1e-3)
    adaptive_stats = dict(adaptive.last_stats) if cfg["method"] == "dopri5" else None
    adjoint_stats = dict(adaptive.last_adjoint_stats) if cfg["method"] == "dopri5" else None

    # ---- per-kernel timing: the same K steps again with every kernel of the solve bracketed by CUDA events on the launching
    # stream.  A separate pass because an event between two kernels removes their programmatic-dependent-launch overlap,
    # which would slow the timed region above; the per-kernel durations are what the roofline uses.  These are SERIALISED
    # (no-overlap) times: their sum may exceed ms_per_step. ----
    _capi.profile_enable(list(_capi.PROF_CLASSES))
    barrier()
    pe = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    pe[0].record()
    for _ in range(args.steps):
        flush.zero_()
        step(src, static, labels)
    pe[1].record()
    barrier()
    prof = _capi.profile_read()
    _capi.profile_enable([])
    profiled_ms_per_step = pe[0].elapsed_time(pe[1]) / args.steps

    # ---- end to end: pinned host inputs -> H2D -> step -> loss D2H, wall clock, max over ranks ----
    for _ in range(2):
        step(src_h.to(dev, non_blocking=True), static_h.to(dev, non_blocking=True),
             labels_h.to(dev, non_blocking=True)).item()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c = src_h.to(dev, non_blocking=True)
        s = static_h.to(dev, non_blocking=True)
        y = labels_h.to(dev, non_blocking=True)
        loss_value = step(c, s, y).item()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = units_per_step / float(te.item())
    h2d = src_h.numel() * 4 + static_h.numel() * 4 + labels_h.numel() * labels_h.element_size()
    d2h = 4

    # ---- the other tensor-core mode on the same workload (same timing rules), reported next to the headline ----
    alt = None
    if args.precision in ("bf16x3", "bf16") and cfg["method"] == "rk4" and not args.no_alt_mode:
        alt_prec = "bf16" if args.precision == "bf16x3" else "bf16x3"
        if alt_prec in solver._PRECISIONS:
            model.precision = alt_prec
            for _ in range(3):
                step(src, static, labels)
            barrier()
            ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            for a2, b2 in ev2:
                flush.zero_()
                a2.record()
                step(src, static, labels)
                b2.record()
            barrier()
            t2 = torch.tensor([sum(a2.elapsed_time(b2) for a2, b2 in ev2)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t2, op=dist.ReduceOp.MAX)
            alt_ms = float(t2.item()) / args.steps
            alt = {alt_prec: {"value": units_per_step / (alt_ms * 1e-3), "unit": UNIT, "ms_per_step": alt_ms,
                              "dtype": DTYPES[alt_prec], "parity_bound": PARITY_BOUNDS[alt_prec]}}
            model.precision = args.precision

    if rank == 0:
        peaks = measured_peaks()
        kernel_ms = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps} for k, v in prof.items()
                     if v[1]}
        n_stage_total = (K - 1) * (4 if cfg["method"] == "rk4" else 7)
        nsteps = K - 1
        roof = roofline(cfg, args, B, prof, peaks, nsteps, profiled_ms_per_step)
        # HBM streams of the path (north_star: achieved GB/s for the path-derivative and state streams).  dx_all: one launch per
        # solve gathers dX/dt for every stage (reads B*C*4 per stage from derivs, writes B*Cp*4).
        C = path_channels(cfg)
        Cp = -(-C // 8) * 8 if args.precision != "fp32" else -(-C // 4) * 4
        dx_bytes = n_stage_total * B * (C + Cp) * 4
        dx_ms, dx_n = prof.get("other", (0.0, 0))
        hbm_streams = {
            "peak_GBps": peaks.get("hbm_gbs"),
            "dx_all": {"algorithmic_bytes_per_launch": dx_bytes, "avg_launch_us": dx_ms / max(dx_n, 1) * 1e3,
                       "achieved_GBps": (dx_bytes / (dx_ms / max(dx_n, 1) * 1e-3) / 1e9) if dx_n else None,
                       "timing": "CUDA events around the launch, same pass as roofline"}}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": DTYPES[args.precision], "parity_bound": PARITY_BOUNDS[args.precision],
            "data": "synthetic",
            "config": config_dict(cfg, B, world, args.scaling),
            "precision": args.precision,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(te.item()) * 1e3, "loss": loss_value},
            "gpu_launches": gpu_launches,
            "clocks": clocks,
            "roofline": roof,
            "kernel_ms": kernel_ms,
            "kernel_ms_note": "serialised per-kernel CUDA-event times of a second pass (%.1f ms/step with the events in place)"
                              % profiled_ms_per_step,
            "hbm_streams": hbm_streams,
            "algorithmic_tflops": (12.0 if cfg["method"] == "rk4" else 0.0) * eval_flops_per_sample(cfg) * units_per_step /
                                  (ms_per_step * 1e-3) / 1e12,
            "samples_per_sec": world * B / (ms_per_step * 1e-3),
        }
        if alt:
            line["other_modes"] = alt
        if adaptive_stats:
            line["dopri5"] = {"forward": {k: adaptive_stats.get(k) for k in ("attempted", "accepted", "nfe")},
                              "adjoint": {k: adjoint_stats.get(k) for k in ("attempted", "accepted", "nfe")}}
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count()
            torch.set_num_threads(cores)
            v, secs, reps, kind, note = time_cpu(cfg, B, args.cpu_baseline_seconds)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": "%s; %d of %d series, all %d knot intervals, fwd+bwd+Adam, best of %d steps, "
                                              "%.1f s per step" % (note, B, B, K - 1, reps, secs)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def roofline(cfg, args, B, prof, peaks, nsteps, profiled_ms_per_step):
    """Roofline of the dominant kernel from the live CUDA-event times of the profiling pass.  Persistent solve kernels (one
    launch per pass): the backward launch, algorithmic FLOPs = dgrad + wgrad of every vector-field evaluation of the pass
    (2x the forward figure; the recompute of the pre-activations is not credited).  Per-stage launches: the
    final-layer backward kernel."""
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    tj = json.load(open(tpath)) if os.path.exists(tpath) else {}
    n_eval = nsteps * (4 if cfg["method"] == "rk4" else 7)
    if prof.get("solve_bwd", (0, 0))[1]:
        ms, n = prof["solve_bwd"]
        flops = 2.0 * eval_flops_per_sample(cfg) * B * n_eval
        name = "persist_bwd_kernel (whole backward pass of one solve: recompute + dgrad + wgrad of every RK stage)"
        traffic = tj.get(args.precision, {}).get("persist_bwd_dram_bytes_per_launch")
        extra = {}
        if prof.get("solve_fwd", (0, 0))[1]:
            fms, fn = prof["solve_fwd"]
            fl = eval_flops_per_sample(cfg) * B * n_eval
            extra["persist_fwd"] = {"achieved": fl / (fms / fn * 1e-3) / 1e12, "avg_launch_us": fms / fn * 1e3,
                                    "flops_per_launch": fl}
    elif prof.get("field_bwd", (0, 0))[1]:
        ms, n = prof["field_bwd"]
        flops = 2.0 * field_flops_per_sample(cfg) * B
        name = ("tc_field_bwd_kernel" if args.precision != "fp32" else "field_bwd_kernel") + \
            " (final-layer recompute + dgrad + wgrad of one RK stage)"
        traffic = tj.get(args.precision, {}).get("field_bwd_dram_bytes_per_launch")
        extra = {}
        if prof.get("field_fwd", (0, 0))[1]:
            fms, fn = prof["field_fwd"]
            fl = field_flops_per_sample(cfg) * B
            extra["field_fwd"] = {"achieved": fl / (fms / fn * 1e-3) / 1e12, "avg_launch_us": fms / fn * 1e3,
                                  "flops_per_launch": fl}
    else:
        return None
    avg = ms / n * 1e-3
    achieved = flops / avg / 1e12
    r = {"kernel": name, "bound": "tensor", "achieved": achieved, "peak": peaks["tensor_tflops"], "unit": "TFLOP/s",
         "frac": achieved / peaks["tensor_tflops"], "traffic": traffic, "peak_source": peaks["source"],
         "flops_per_launch": flops, "avg_launch_us": avg * 1e6,
         "timing": "CUDA events around every launch, second pass of the same steps (%.1f ms/step with the events in place)"
                   % profiled_ms_per_step}
    r.update(extra)
    return r


def run_check(args, cfg, rank, world, dev):
    """N-rank all-reduced (averaged) gradients of the sharded batch == single-rank gradients of the concatenated batch, fp32
    path, 1e-5 relative, over NCCL.  Every rank builds the full batch; rank r solves rows [r*B, (r+1)*B)."""
    import torch.distributed as dist
    import ncde_b200
    import torchcde_b200 as tc
    from torchcde_b200.distributed import allreduce_gradients
    B = min(args.batch_per_gpu, 128)
    torch.manual_seed(1)

    def make():
        torch.manual_seed(1)
        return ncde_b200.NeuralCDE(path_channels(cfg), cfg["H"], cfg["out"], static_dim=cfg["S"] or None,
                                   hidden_hidden_dim=cfg["HH"], num_layers=cfg["n_layers"], interpolation=cfg["interp"],
                                   adjoint=False, solver="rk4", return_sequences=cfg["online"], precision="fp32").to(dev)

    x, static, labels = synth_batch(B * world, seed=5, cfg=cfg)
    coeffs = tc.linear_interpolation_coeffs(x.to(dev), rectilinear=0 if cfg["interp"] == "rectilinear" else None)
    static, labels = static.to(dev), labels.to(dev)

    def grads(model, sl):
        out = model((static[sl], coeffs[sl].contiguous())) if cfg["S"] else model(coeffs[sl].contiguous())
        # sum-reduction so that the average over ranks of per-shard means equals the full-batch mean
        loss_fn(cfg, out, labels[sl]).backward()
        return [p.grad for p in model.parameters()]

    m_full = make()
    g_full = grads(m_full, slice(0, B * world))
    m_shard = make()
    grads(m_shard, slice(rank * B, (rank + 1) * B))
    if world > 1:
        allreduce_gradients(m_shard.parameters(), average=True)
    worst = 0.0
    for gf, p in zip(g_full, m_shard.parameters()):
        worst = max(worst, float((gf - p.grad).abs().max() / gf.abs().max().clamp_min(1e-30)))
    t = torch.tensor([worst], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ok = float(t.item()) <= 1e-5
        print(json.dumps({"check": "allreduced_gradients_equal_full_batch", "n_gpus": world, "rows_per_rank": B,
                          "max_rel_err": float(t.item()), "bound": 1e-5, "ok": ok, "backend": "nccl" if world > 1 else "none"}),
              flush=True)
        assert ok


if __name__ == "__main__":
    main()
