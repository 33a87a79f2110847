"""Shared helpers of the GPU parity tests: the oracle run with its ReLU margins recorded, and the rule that decides which
batch rows may be excluded from a 1e-5 gradient comparison.

Why rows may be excluded at all: the reference's vector field has ReLU hidden layers.  A pre-activation that sits within
rounding distance of zero takes a different branch in two fp32 implementations that sum a dot product in different
orders; the hidden state barely moves (the unit's output is ~0 either way) but the gradient of that row changes by the
whole contribution of the unit.  The reference's own fp32 gradient differs from its fp64 gradient by 1e-4..1e-3 on such
rows (tools/diag_rows.py).  The rule below ties every excluded row to an ORACLE-side fact instead of to the GPU's error:

  * a row may be excluded only if, in the oracle's fp32 run, some hidden pre-activation of that row came closer to zero
    than `margin_tol` = 32 * (measured relative state error, floored at 2^-22) * (largest |pre-activation| of the run)
    — the distance two correct fp32 implementations can disagree by;
  * excluded rows must be rare (<= 2 % of the batch, at least 1 allowed);
  * against the oracle run in fp64, the GPU must not have more rows beyond 1e-5 than the reference's own fp32 arithmetic
    has, up to a factor 2 (+2 rows): it is as close to the exact gradient as the reference is.
"""
import copy

import torch

from oracle import cde_oracle as O


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


class ReluMargins:
    """Forward hooks on every Linear that feeds a ReLU of the oracle's vector field: per batch row, the smallest
    |pre-activation| seen over all evaluations, and the largest magnitude overall."""

    def __init__(self, func):
        self.min_abs = None
        self.max_abs = 0.0
        self.handles = []
        mods = list(func.modules())
        seen = set()
        if hasattr(func, "net_to_hh"):
            seq = list(func.net_to_hh)
            for m, nxt in zip(seq[:-1], seq[1:]):
                if isinstance(m, torch.nn.Linear) and isinstance(nxt, torch.nn.ReLU) and id(m) not in seen:
                    seen.add(id(m))
                    self.handles.append(m.register_forward_hook(self._hook))
        else:   # ToyField: linear0 / linear1 are followed by relu()
            for name in ("linear0", "linear1"):
                self.handles.append(getattr(func, name).register_forward_hook(self._hook))
        assert self.handles, mods

    def _hook(self, module, inp, out):
        a = out.detach().abs()
        row_min = a.amin(dim=-1)
        self.min_abs = row_min if self.min_abs is None else torch.minimum(self.min_abs, row_min)
        self.max_abs = max(self.max_abs, float(a.max()))

    def close(self):
        for h in self.handles:
            h.remove()


def oracle_solve(func, path_kind, coeffs, z0, w, online, method="rk4", step_size=1, dtype=torch.float32, margins=False,
                 row_mask=None):
    """Oracle forward + backward of sum(out * w).  Returns out, grad_z0, {param grads}, ReluMargins or None."""
    f = copy.deepcopy(func).to(dtype)
    Xr = (O.CubicPath if path_kind == "cubic" else O.LinearPath)(coeffs.to(dtype))
    t = Xr.grid_points if online else Xr.interval
    ww = w.to(dtype).clone()
    if row_mask is not None:
        ww[row_mask] = 0
    z0r = z0.clone().to(dtype).requires_grad_(True)
    m = ReluMargins(f) if margins else None
    out = O.cdeint(Xr, f, z0r, t, adjoint=False, method=method, options={"step_size": step_size})
    (out * ww).sum().backward()
    if m is not None:
        m.close()
    return out.detach(), z0r.grad.detach(), {n: p.grad.detach().clone() for n, p in f.named_parameters()}, m


def excluded_rows(gz_gpu, gz_ref32, gz_ref64, out_gpu, out_ref32, margins, tol=1e-5):
    """Rows whose z0-gradient misses `tol`; asserts the rule in the module docstring.  Returns a bool mask (B,)."""
    scale = gz_ref32.abs().max().double()
    row_err = (gz_gpu.double().cpu() - gz_ref32.double()).abs().amax(1) / scale
    bad = row_err > tol
    B = gz_ref32.shape[0]
    assert int(bad.sum()) <= max(1, B // 50), ("too many rows beyond tol", int(bad.sum()), float(row_err.max()))
    if bad.any():
        state_err = max(rel(out_gpu, out_ref32), 2.0 ** -22)
        margin_tol = 32.0 * state_err * margins.max_abs
        row_margin = margins.min_abs.reshape(B)
        loose = bad & ~(row_margin < margin_tol)
        assert not loose.any(), ("rows beyond tol without a ReLU pre-activation near zero in the oracle",
                                 row_margin[loose].tolist(), margin_tol, row_err[loose].tolist())
    if gz_ref64 is not None:
        s64 = gz_ref64.abs().max()
        n_gpu = int(((gz_gpu.double().cpu() - gz_ref64).abs().amax(1) / s64 > tol).sum())
        n_cpu = int(((gz_ref32.double() - gz_ref64).abs().amax(1) / s64 > tol).sum())
        assert n_gpu <= 2 * n_cpu + 2, ("GPU has more ill-conditioned rows than the reference's own fp32 run", n_gpu, n_cpu)
    return bad
