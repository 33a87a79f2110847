#!/usr/bin/env python
"""Benchmark of the Neural-CDE solve hot path (BASELINE.json metric: NCDE fwd+bwd sequence-steps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp32|bf16] [--impl b200|reference]

Workload (config.workload): BASELINE.json configs[4], "MIMIC-IV-shaped synthetic online sepsis": per GPU 1024 series
(8192 over 8 GPUs), 72 hourly steps -> 143 rectilinear knots, 100 channels (time + 99), 5 static features, hidden =
hidden-hidden = 128, 3 vector-field layers, 3/8-rule RK4 with step 1, online outputs, BCE loss on the de-duplicated
outputs, backprop through the solver (adjoint=False, as every config of the reference), Adam step.

One "step" = one forward + backward + optimiser pass over one batch.  seq-steps = B * (K - 1) per step.
Prints ONE JSON line on rank 0.
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

CFG = dict(name="cfg5_mimic_iv_shaped_online_sepsis", B=1024, L=72, C=100, S=5, H=128, HH=128, n_layers=3, out=1,
           obs_rate=0.2)
METRIC = "ncde_fwd_bwd_seq_steps_per_sec"
UNIT = "seq-steps/s"


def synth_batch(B, seed, cfg=CFG):
    """Seeded synthetic batch of the cfg-5 shape: time channel = arange (get_data/common.py:178-184), z-normalised
    values observed with probability obs_rate (NaN elsewhere), first row NaN -> 0 (transformers.py:53-55)."""
    g = torch.Generator().manual_seed(seed)
    L, C = cfg["L"], cfg["C"]
    x = torch.randn(B, L, C, generator=g)
    x[..., 0] = torch.arange(L, dtype=torch.float32)
    miss = torch.rand(B, L, C, generator=g) > cfg["obs_rate"]
    miss[..., 0] = False
    x[miss] = float("nan")
    first = x[:, 0, :]
    first[torch.isnan(first)] = 0.0
    static = torch.randn(B, cfg["S"], generator=g)
    labels = (torch.rand(B, L, generator=g) < 0.1).float()
    return x, static, labels


def field_flops_per_sample(cfg=CFG):
    """Algorithmic FLOPs of the final layer + contraction per vector-field evaluation per sample:
    2*HH*H*C for the GEMM + 2*H*C for the f.dX contraction."""
    return 2.0 * cfg["HH"] * cfg["H"] * cfg["C"] + 2.0 * cfg["H"] * cfg["C"]


def eval_flops_per_sample(cfg=CFG):
    H, HH, C, n = cfg["H"], cfg["HH"], cfg["C"], cfg["n_layers"]
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
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port of the reference's CPU path on the host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_step_fn(B, seed=0):
    from oracle import cde_oracle as O
    torch.manual_seed(1)
    cfg = CFG
    func = O.SharedMLPField(cfg["C"], cfg["H"], cfg["HH"], cfg["n_layers"])
    initial = torch.nn.Linear(cfg["C"] + cfg["S"], cfg["H"])
    readout = torch.nn.Linear(cfg["H"], cfg["out"])
    params = list(func.parameters()) + list(initial.parameters()) + list(readout.parameters())
    opt = torch.optim.Adam(params, lr=1e-3)
    x, static, labels = synth_batch(B, seed)
    coeffs = O.linear_interpolation_coeffs(x, rectilinear=0)
    lossf = torch.nn.BCEWithLogitsLoss()

    def step():
        opt.zero_grad(set_to_none=True)
        out = O.ncde_forward(coeffs, func, initial, readout, "rectilinear", "rk4", False, True, static=static,
                             options={"step_size": 1})
        loss = lossf(out.squeeze(-1), labels)
        loss.backward()
        opt.step()
        return float(loss.detach())

    return step, B * (2 * cfg["L"] - 2)


def time_cpu(B, reps):
    step, units = cpu_reference_step_fn(B)
    tiny, _ = cpu_reference_step_fn(4)
    tiny()  # lazy initialisation outside the timing
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        step()
        t.append(time.perf_counter() - t0)
    return units / min(t), min(t)


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    B = args.ref_batch
    step, units = cpu_reference_step_fn(B)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = units / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": CFG["name"], "batch_per_gpu": CFG["B"], "knots": 2 * CFG["L"] - 1,
                       "channels": CFG["C"], "hidden": CFG["H"], "solver": "rk4(3/8) step 1, backprop"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "oracle/cde_oracle.py (PyTorch CPU restatement of the reference path), "
                                       "%d of %d series per step, all %d steps" % (B, CFG["B"], 2 * CFG["L"] - 2)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("NCDE_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--batch-per-gpu", type=int, default=CFG["B"])
    ap.add_argument("--ref-batch", type=int, default=128, help="series per step of the CPU reference arm")
    ap.add_argument("--cpu-baseline-batch", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"

    import torch.distributed as dist
    import torchcde_b200 as tc
    import ncde_b200
    from torchcde_b200 import _capi, solver
    from torchcde_b200.distributed import allreduce_gradients

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = CFG
    B = args.batch_per_gpu
    K = 2 * cfg["L"] - 1

    # model: identical initial weights on every rank
    torch.manual_seed(1)
    model = ncde_b200.NeuralCDE(cfg["C"], cfg["H"], cfg["out"], static_dim=cfg["S"], hidden_hidden_dim=cfg["HH"],
                                num_layers=cfg["n_layers"], interpolation="rectilinear", adjoint=False, solver="rk4",
                                return_sequences=True, precision=args.precision).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    lossf = torch.nn.BCEWithLogitsLoss()

    # data: this rank's shard of the global batch; coefficients are built once on the device (offline in the reference)
    x, static_h, labels_h = synth_batch(B, seed=100 + rank)
    coeffs = tc.linear_interpolation_coeffs(x.to(dev), rectilinear=0)
    static, labels = static_h.to(dev), labels_h.to(dev)
    coeffs_h = coeffs.cpu().pin_memory()
    static_h, labels_h = static_h.pin_memory(), labels_h.pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    launches = {"n": 0}

    def step(c, s, y):
        opt.zero_grad(set_to_none=True)
        out = model((s, c))
        loss = lossf(out.squeeze(-1), y)
        loss.backward()
        if world > 1:
            allreduce_gradients(model.parameters(), average=True)
        opt.step()
        # kernels of libncde_b200 this step: solve fwd + bwd, linear_derivs (spline ctor), path_eval (X(0))
        launches["n"] += solver.last_launches["fwd"] + solver.last_launches["bwd"] + 2
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(coeffs, static, labels)
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
        step(coeffs, static, labels)
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

    # ---- per-kernel timing: the same K steps again with every stage kernel bracketed by CUDA events on the launching
    # stream.  A separate pass because an event between two kernels removes their programmatic-dependent-launch overlap,
    # which would slow the timed region above; the per-kernel durations are what the roofline uses. ----
    _capi.profile_enable(["field_fwd", "field_bwd", "hidden_fwd", "hidden_bwd", "hidden_wgrad", "other"])
    barrier()
    pe = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    pe[0].record()
    for _ in range(args.steps):
        flush.zero_()
        step(coeffs, static, labels)
    pe[1].record()
    barrier()
    prof = _capi.profile_read()
    _capi.profile_enable([])
    profiled_ms_per_step = pe[0].elapsed_time(pe[1]) / args.steps

    # ---- end to end: pinned host inputs -> H2D -> step -> loss D2H, wall clock, max over ranks ----
    for _ in range(2):
        step(coeffs_h.to(dev, non_blocking=True), static_h.to(dev, non_blocking=True),
             labels_h.to(dev, non_blocking=True)).item()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c = coeffs_h.to(dev, non_blocking=True)
        s = static_h.to(dev, non_blocking=True)
        y = labels_h.to(dev, non_blocking=True)
        loss_value = step(c, s, y).item()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = units_per_step / float(te.item())
    h2d = coeffs_h.numel() * 4 + static_h.numel() * 4 + labels_h.numel() * 4
    d2h = 4

    if rank == 0:
        peaks = measured_peaks()
        bwd_ms, bwd_n = prof["field_bwd"]
        fwd_ms, fwd_n = prof["field_fwd"]
        # algorithmic FLOPs per field_bwd launch: dgrad + wgrad of the final layer and contraction for B rows
        # (= 2x the forward figure; the tanh recompute is not credited)
        flops_bwd = 2.0 * field_flops_per_sample() * B
        flops_fwd = field_flops_per_sample() * B
        avg_bwd = bwd_ms / max(bwd_n, 1) * 1e-3
        avg_fwd = fwd_ms / max(fwd_n, 1) * 1e-3
        achieved = flops_bwd / avg_bwd / 1e12 if bwd_n else None
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(args.precision, {}).get("field_bwd_dram_bytes_per_launch")
        kernel_ms = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps} for k, v in prof.items()
                     if v[1]}
        # HBM streams of the path (north_star: achieved GB/s for the path-derivative and state streams).  dx_all: one launch per
        # solve gathers dX/dt for every stage (reads B*C*4 per stage from derivs, writes B*Cp*4); records: what the forward pass
        # saves per stage for the backward pass and the backward pass reads back.
        n_stage_total = (K - 1) * 4
        Cp = -(-cfg["C"] // (8 if args.precision == "bf16" else 4)) * (8 if args.precision == "bf16" else 4)
        dx_bytes = n_stage_total * B * (cfg["C"] + Cp) * 4
        dx_ms, dx_n = prof.get("other", (0.0, 0))
        rec_bytes_per_series = ((cfg["n_layers"] + 1) * 256 + 4 * Cp) if args.precision == "bf16" else \
            4 * (cfg["H"] + cfg["n_layers"] * cfg["HH"] + Cp)
        rec_bytes = 2 * n_stage_total * B * rec_bytes_per_series   # written forward, read backward
        hbm_streams = {
            "peak_GBps": peaks.get("hbm_gbs"),
            "dx_all": {"algorithmic_bytes_per_launch": dx_bytes, "avg_launch_us": dx_ms / max(dx_n, 1) * 1e3,
                       "achieved_GBps": (dx_bytes / (dx_ms / max(dx_n, 1) * 1e-3) / 1e9) if dx_n else None,
                       "timing": "CUDA events around the launch, same pass as roofline"},
            "stage_records": {"bytes_per_step": rec_bytes, "GBps_over_step": rec_bytes / (ms_per_step * 1e-3) / 1e9,
                              "note": "spread over the whole sequential step: never the bound"}}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "bf16 tiles, f32 accumulate/state",
            "data": "synthetic",
            "config": {"workload": cfg["name"], "batch_per_gpu": B, "global_batch": B * world, "length": cfg["L"],
                       "knots": K, "channels": cfg["C"], "static": cfg["S"], "hidden": cfg["H"],
                       "hidden_hidden": cfg["HH"], "vector_field_layers": cfg["n_layers"],
                       "solver": "rk4(3/8) step 1, backprop through solver, Adam", "precision": args.precision,
                       "l2": "256 MB flush between timed steps; per-step working set 1.5 GB >> 126 MB L2"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(te.item()) * 1e3, "loss": loss_value},
            "gpu_launches": gpu_launches,
            "clocks": clocks,
            "roofline": {"kernel": ("tc_field_bwd_kernel<8>" if args.precision == "bf16" else "field_bwd_kernel") +
                                   " (final-layer recompute + dgrad + wgrad of one RK stage)", "bound": "tensor",
                         "achieved": achieved, "peak": peaks["tensor_tflops"], "unit": "TFLOP/s",
                         "frac": (achieved / peaks["tensor_tflops"]) if achieved else None, "traffic": traffic,
                         "peak_source": peaks["source"], "flops_per_launch": flops_bwd,
                         "avg_launch_us": avg_bwd * 1e6,
                         "timing": "CUDA events around every launch, second pass of the same %d steps "
                                   "(%.1f ms/step with the events in place)" % (args.steps, profiled_ms_per_step),
                         "field_fwd": {"achieved": flops_fwd / avg_fwd / 1e12 if fwd_n else None,
                                       "avg_launch_us": avg_fwd * 1e6, "flops_per_launch": flops_fwd}},
            "kernel_ms": kernel_ms,
            "hbm_streams": hbm_streams,
            "algorithmic_tflops": 12.0 * eval_flops_per_sample() * units_per_step / (ms_per_step * 1e-3) / 1e12,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count()
            torch.set_num_threads(cores)
            v, secs = time_cpu(args.cpu_baseline_batch, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "oracle port, %d of %d series, all %d steps, fwd+bwd+Adam, %.1f s"
                                              % (args.cpu_baseline_batch, B, K - 1, secs)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
