"""Throughput of the ragged-batch preprocessing on a cfg-5-sized raw set, with the per-series CPU loop of the
reference (oracle restatement of get_data/transformers.py:50-85) timed beside it on a bounded sample.

    python tools/bench_ragged.py            -> one JSON line per method

Algorithmic bytes per launch: the raw set read once (n*Lmax*C*4) and the coefficients written once (linear: same size;
rectilinear: 2x; cubic: 4x): HBM-bound copy/scan work.  NCDE_RAGGED_NO_STAGING=1 selects the un-staged kernels (A/B).
"""
import json, os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "online-neural-cdes_b200")]
import torch
from ncde_b200 import preprocessing as P
from oracle import cde_oracle as O

peaks = {}
try:
    peaks = json.load(open(os.path.join(R, "MEASURED_PEAKS.json")))
except Exception:
    pass
hbm = None
for k, v in peaks.items():
    if "hbm" in k.lower() and isinstance(v, (int, float)):
        hbm = float(v)
g = torch.Generator().manual_seed(8)
n, Lmax, C = 8192, 72, 100
lengths = torch.randint(2, Lmax + 1, (n,), generator=g)
x = torch.randn(n, Lmax, C, generator=g)
x[..., 0] = torch.arange(Lmax, dtype=torch.float32)
drop = torch.rand(n, Lmax, C, generator=g) < 0.75
drop[..., 0] = False
x[drop] = float("nan")
for i in range(n):
    x[i, int(lengths[i]):] = float("nan")
xd = x.cuda()
series_dev = [xd[i, :int(lengths[i])] for i in range(n)]
for method in ("linear", "rectilinear", "cubic"):
    # device-resident packed input: time the library call itself (pack once outside)
    from torchcde_b200 import _capi
    code = P._METHODS[method]
    L_ = _capi.lib()
    len_dev = lengths.to(torch.int32).cuda()
    Kmax = P._out_rows(code, Lmax)
    Cout = 4 * C if method == "cubic" else C
    out = torch.empty(n, Kmax, Cout, device="cuda")
    scratch = torch.empty(L_.ncde_ragged_scratch_bytes(code, 0, n, Lmax, C), dtype=torch.uint8, device="cuda")
    flags = torch.zeros(1, dtype=torch.int32, device="cuda")
    def call():
        _capi.check(L_.ncde_ragged_interpolate(code, 0, xd.data_ptr(), len_dev.data_ptr(), out.data_ptr(), n, Lmax, C, 0, 1, 0, 1,
                                               scratch.data_ptr(), flags.data_ptr(), _capi.stream_ptr(xd.device)))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        call()
    ts = []
    for _ in range(5):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); call(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    raw_bytes = n * Lmax * C * 4
    alg = raw_bytes + out.numel() * 4          # the raw set read once, the coefficients written once
    # reference-style CPU loop on a bounded sample
    m = 256
    sample = [x[i, :int(lengths[i])].clone() for i in range(m)]
    t0 = time.perf_counter()
    O.interpolation_transform(sample, method)
    cpu_s = time.perf_counter() - t0
    line = {"metric": "ragged_interpolation_series_per_sec", "method": method, "value": n / (ms / 1e3), "unit": "series/s",
            "ms_per_launch_set": ms, "staged": os.environ.get("NCDE_RAGGED_NO_STAGING") is None, "config": {"workload": "cfg5_raw_set", "series": n, "max_length": Lmax, "channels": C,
                                                "missing": 0.75, "l2": "256 MB flush between timed calls"},
            "roofline": {"bound": "hbm", "achieved": alg / (ms / 1e3) / 1e9, "peak": hbm, "unit": "GB/s",
                         "frac": (alg / (ms / 1e3) / 1e9 / hbm) if hbm else None, "algorithmic_bytes": alg},
            "cpu_baseline": {"value": m / cpu_s, "unit": "series/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": "oracle loop over %d of %d series, %.2f s" % (m, n, cpu_s)}}
    print(json.dumps(line))
