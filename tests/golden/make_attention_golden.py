"""Mints tests/golden/attention.pt from the REAL reference src/ncde/attention.py (run in the build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_attention_golden.py

`autots` (un-vendored, requirements.txt:2) is stubbed with the three classes attention.py uses, given the behaviour their use in
the reference implies (experiments/ingredients/loader.py:190-196 batches coefficient lists the same way): PadRaggedTensors pads a
list of (length_i, channels) tensors with NaN to the longest, ForwardFill fills NaN with the last value along time, SimplePipeline
chains transforms.  Cases keep every attention weight at a margin from the 1 / length threshold so that the kept set is not a
rounding decision.
"""
import os
import sys
import types
import warnings

import torch

REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(REF, "modules/torchcde"))
sys.path.insert(0, os.path.join(REF, "modules/torchdiffeq"))
sys.path.insert(0, REF)


class PadRaggedTensors:
    def transform(self, data):
        if isinstance(data, torch.Tensor):
            return data
        n = max(d.size(0) for d in data)
        return torch.stack([torch.cat([d, torch.full((n - d.size(0),) + tuple(d.shape[1:]), float("nan"), dtype=d.dtype)]) for d in data])


class ForwardFill:
    def transform(self, x):
        import torchcde
        return torchcde.misc.forward_fill(x) if hasattr(torchcde, "misc") else x


class SimplePipeline:
    def __init__(self, steps):
        self.steps = steps

    def transform(self, x):
        for s in self.steps:
            x = s.transform(x)
        return x


_a = types.ModuleType("autots")
_p = types.ModuleType("autots.preprocessing")
_p.ForwardFill, _p.PadRaggedTensors, _p.SimplePipeline = ForwardFill, PadRaggedTensors, SimplePipeline
_a.preprocessing = _p
sys.modules["autots"] = _a
sys.modules["autots.preprocessing"] = _p
import torchcde  # noqa: E402
import torchcde.misc  # noqa: E402,F401
import src.ncde.sparsemax as _sm  # noqa: E402
_sm.device = torch.device("cpu")
from src.ncde.attention import AttentionNeuralCDE  # noqa: E402

g = torch.Generator().manual_seed(97531)
out = {}
cases = {
    "softmax_backprop": dict(adjoint=False, sparsemax=False, static_dim=None),
    "sparsemax_backprop_static": dict(adjoint=False, sparsemax=True, static_dim=3),
    "softmax_adjoint_forwards": dict(adjoint=True, sparsemax=False, static_dim=None, run_backwards=False),
}
for name, kw in cases.items():
    for seed in range(50):
        torch.manual_seed(100 + seed)
        B, L, C, H, O_ = 6, 9, 3, 5, 2
        x = torch.randn(B, L, C, generator=g)
        x[..., 0] = torch.arange(L, dtype=torch.float32)
        x[..., 1:] = x[..., 1:].cumsum(-2) * 0.5
        coeffs = torchcde.linear_interpolation_coeffs(x)
        model = AttentionNeuralCDE(C, H, O_, **kw)
        with torch.no_grad():   # make the attention weights vary over time enough for a clear kept set
            for p_ in model.attention[1].parameters():
                p_.mul_(3.0)
        static = torch.randn(B, kw["static_dim"], generator=g) if kw.get("static_dim") else None
        inputs = coeffs if static is None else [static, coeffs]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            hidden = model.encoder(inputs)
            att = model.attention(hidden if static is None else [static, hidden.clone()])
        margin = float((att - 1.0 / L).abs().min())
        kept = (att > 1.0 / L).sum(1).flatten()
        if margin > 2e-4 and int(kept.min()) >= 2 and int(kept.max()) < L:
            break
    else:
        raise SystemExit("no well-separated case found for " + name)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        y = model(coeffs if static is None else [static, coeffs])
    w = torch.randn(y.shape, generator=g)
    (y * w).sum().backward()
    out[name] = {"kwargs": kw, "dims": (C, H, O_), "coeffs": coeffs, "static": static, "w": w, "out": y.detach().clone(),
                 "attention": att.detach().clone(), "kept": kept.clone(),
                 "state_dict": {k: v.clone() for k, v in model.state_dict().items()},
                 "grads": {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}}
    print(name, tuple(y.shape), "kept", kept.tolist(), "margin %.1e" % margin, len(out[name]["grads"]), "grads")
# sparsemax known answers from the reference module itself
sm = _sm.Sparsemax(dim=1)
zin = torch.randn(4, 7, 1, generator=g) * 2
out["sparsemax"] = {"in": zin, "out": sm(zin).clone()}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "attention.pt")
torch.save(out, path)
print("wrote", path, os.path.getsize(path))
