"""Linear / rectilinear hybrid path construction on the GPU (mirrors src/ncde/interpolation.py:186-253 of the reference).

Regularly sampled channels are interpolated linearly, sparsely sampled ones rectilinearly, and a knot is kept only where
a rectilinear channel (or time) changes — a much shorter path than the full rectilinear one for sparse ICU measurements.
Every step is copy / select work, so the result is bit-identical to the reference (its known-answer test is
tests/test_hybrid.py).  SmoothLinearInterpolation (the cubic / quintic gradient-matching regions of the same file) is not
implemented yet.
"""
import torch

import torchcde_b200 as torchcde
from torchcde_b200 import _capi


def _prepare_linear_rectilinear_hybrid(data, rectilinear_indices, time_index=0):
    """data: (batch, length, channels), NaN = missing; mutated in place exactly like the reference (the linear channels are
    filled, first-row NaNs become 0).  Returns (batch, kept_length, channels)."""
    assert isinstance(rectilinear_indices, list)
    _capi.require_cuda(data)
    assert data.dim() == 3, "the hybrid scheme takes exactly one batch dimension (as the reference does)"

    # interpolation.py:208-216 — first linearly interpolate the non rectilinear indices
    time_and_rect_indices = [time_index] + rectilinear_indices
    non_rect_indices = [x for x in range(data.size(-1)) if x not in time_and_rect_indices]
    data[..., non_rect_indices] = torchcde.linear_interpolation_coeffs(data[..., non_rect_indices], initial_value_if_nan=0.0)

    # :218-221 — then rectilinear everything (time index 0, as hard-coded there)
    full_rectilinear = torchcde.linear_interpolation_coeffs(data, rectilinear=0, initial_value_if_nan=0.0).contiguous()

    # :223-253 — shift the linear channels, drop the rows where nothing changed, pad + forward fill: one kernel
    B, K, C = full_rectilinear.shape
    kind = torch.zeros(C, dtype=torch.int32)
    kind[time_and_rect_indices] = 1
    kind = kind.to(data.device)
    out = torch.empty_like(full_rectilinear)
    counts = torch.empty(B, dtype=torch.int32, device=data.device)
    _capi.check(_capi.lib().ncde_hybrid_compact(_capi.dtype_code(full_rectilinear), full_rectilinear.data_ptr(), kind.data_ptr(),
                                                out.data_ptr(), counts.data_ptr(), B, K, C, _capi.stream_ptr(data.device)))
    longest = int(counts.max().item()) if B > 0 else 0   # the one synchronisation (the reference's pad_sequence needs it too)
    return out[:, :longest].contiguous()
