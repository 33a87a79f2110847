"""Mints tests/golden/ragged.pt (run in the build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_ragged_golden.py

* `transform`: the REAL reference transformer get_data/transformers.py:7-85 (`Interpolation`, imported by path) on a list of
  series of different lengths with missing values, every method.
* `loader`: experiments/ingredients/loader.py cannot be imported here (sacred, ignite, autots are absent), so its
  rectilinear-intensity block (:100-113) and its per-batch padding (:181-202, autots PadRaggedTensors + ForwardFill) are
  restated below statement by statement with the reference's own torchcde ops; parity for these two is "restated", not "run".
"""
import importlib.util
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(REF, "modules/torchcde"))
sys.path.insert(0, os.path.join(REF, "modules/torchdiffeq"))
import torchcde  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_transformers", os.path.join(REF, "get_data/transformers.py"))
tr = importlib.util.module_from_spec(spec)
spec.loader.exec_module(tr)

g = torch.Generator().manual_seed(97531)


def make_series(n, C, lo, hi, p, dtype=torch.float32):
    out = []
    for _ in range(n):
        L = int(torch.randint(lo, hi + 1, (1,), generator=g))
        x = torch.randn(L, C, generator=g, dtype=torch.float64).to(dtype)
        x[:, 0] = torch.arange(L, dtype=dtype)
        drop = torch.rand(L, C, generator=g) < p
        drop[:, 0] = False
        x[drop] = float("nan")
        if L > 3 and torch.rand(1, generator=g) < 0.3:
            x[:, 1 + int(torch.randint(0, C - 1, (1,), generator=g))] = float("nan")   # a channel never observed
        out.append(x)
    return out


res = {"transform": {}, "loader": {}}
for name, (n, C, lo, hi, p, dtype) in {"mimic_like": (23, 6, 2, 19, 0.6, torch.float32),
                                       "dense": (7, 3, 4, 9, 0.0, torch.float32),
                                       "f64": (9, 4, 3, 12, 0.4, torch.float64)}.items():
    raw = make_series(n, C, lo, hi, p, dtype)
    rec = {"raw": [x.clone() for x in raw]}
    for method in ("linear", "rectilinear", "cubic", "linear_forward_fill"):
        data = [x.clone() for x in raw]
        out = tr.Interpolation(method=method).fit_transform(data)
        rec[method] = [o.clone() for o in out]
        rec["mutated_" + method] = [d.clone() for d in data]
    res["transform"][name] = rec
    print(name, [tuple(o.shape) for o in rec["cubic"][:4]])

# tensor input (equal lengths)
x = torch.randn(5, 8, 4, generator=g)
x[..., 0] = torch.arange(8, dtype=torch.float32)
drop = torch.rand(x.shape, generator=g) < 0.4
drop[..., 0] = False
x[drop] = float("nan")
res["transform"]["tensor"] = {"raw": x.clone(), **{m: tr.Interpolation(method=m).fit_transform(x.clone()) for m in
                                                   ("linear", "rectilinear", "cubic")}}

# loader.py:100-113 (rectilinear-intensity) and :181-202 (sorted, padded batches), restated
raw = make_series(17, 5, 2, 14, 0.5)
data = [x.clone() for x in raw]
temporal = [o.numpy().astype(np.float32) for o in tr.Interpolation(method="rectilinear").fit_transform(data)]
raw_after = data   # `temporal_data_raw` is the list the pipeline zeroed in place (get_data/common.py:108-113)
with_int = []
for i in range(len(temporal)):
    tdata = torch.tensor(np.copy(raw_after[i].numpy()))
    tdata[0, :][tdata[0, :] == 0] = float("nan")
    intensity_cumsum = (~tdata[:, 1:].isnan()).cumsum(axis=0).repeat_interleave(2, 0)
    intensity_cumsum = intensity_cumsum[:-1]
    with_int.append(torch.tensor(np.concatenate([temporal[i], intensity_cumsum.numpy().astype(temporal[i].dtype)], axis=1)))
lengths = [len(x) for x in with_int]
order = sorted(range(len(lengths)), key=lambda k: lengths[k])
sorted_coeffs = [with_int[i] for i in order]
batch_size = 5
batches = []
for i in range(0, len(sorted_coeffs), batch_size):
    chunk = sorted_coeffs[i:i + batch_size]
    padded = torch.nn.utils.rnn.pad_sequence(chunk, batch_first=True, padding_value=float("nan"))   # PadRaggedTensors
    batches.append(torchcde.misc.forward_fill(padded))                                              # ForwardFill
res["loader"] = {"raw": [x.clone() for x in raw], "with_intensity": with_int, "order": order, "batch_size": batch_size,
                 "batches": batches}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ragged.pt")
torch.save(res, path)
print("wrote", path, os.path.getsize(path))
