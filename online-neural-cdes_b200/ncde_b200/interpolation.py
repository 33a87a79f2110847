"""Linear / rectilinear hybrid path construction on the GPU (mirrors src/ncde/interpolation.py:186-253 of the reference).

Regularly sampled channels are interpolated linearly, sparsely sampled ones rectilinearly, and a knot is kept only where
a rectilinear channel (or time) changes — a much shorter path than the full rectilinear one for sparse ICU measurements.
Every step is copy / select work, so the result is bit-identical to the reference (its known-answer test is
tests/test_hybrid.py).

SmoothLinearInterpolation (the same file, :6-183): linear interpolation whose kinks are rounded off by a cubic (first
derivative matched) or quintic (second derivative too) on [t_k, t_k + eps) after every interior knot.  A subclass of the
drop-in LinearInterpolation, so `cdeint` takes it like any other path; the matching coefficients come from
`ncde_smooth_matching_coeffs`, evaluation from `ncde_path_eval_smooth`, and the fused solve reads the same coefficients
through `ncde_path_t.match`.
"""
import torch

import torchcde_b200 as torchcde
from torchcde_b200 import _capi
from torchcde_b200.interpolation_linear import LinearInterpolation, _flatten


class SmoothLinearInterpolation(LinearInterpolation):
    """Arguments as src/ncde/interpolation.py:9-29: `gradient_matching_eps` in (0, 1] switches the matching regions on,
    `match_second_derivatives` selects quintic instead of cubic pieces; `t` may only be given without matching."""

    def __init__(self, coeffs, t=None, gradient_matching_eps=None, match_second_derivatives=False, **kwargs):
        if t is not None:
            assert gradient_matching_eps is None, "times not implemented for gradient_matching_eps"
        super(SmoothLinearInterpolation, self).__init__(coeffs, t=t, **kwargs)
        self.gradient_matching_eps = gradient_matching_eps
        self.match_second_derivatives = match_second_derivatives
        self.match_terms = 6 if match_second_derivatives else 4
        if gradient_matching_eps is not None:
            assert 0 < gradient_matching_eps <= 1
            assert coeffs.dim() == 3, "gradient matching takes exactly one batch dimension (as the reference does)"
            c = coeffs.detach().contiguous()
            n, K, C = _flatten(c)
            match = torch.empty(*c.shape[:-2], K - 2, C, self.match_terms, dtype=c.dtype, device=c.device)
            _capi.check(_capi.lib().ncde_smooth_matching_coeffs(_capi.dtype_code(c), c.data_ptr(), match.data_ptr(), n, K, C,
                                                                float(gradient_matching_eps), self.match_terms,
                                                                _capi.stream_ptr(c.device)))
            self.register_buffer("gradient_matching_coeffs", match)

    def __len__(self):
        return len(self._t_host)

    def _eval(self, t, deriv):
        if self.gradient_matching_eps is None:
            return super(SmoothLinearInterpolation, self)._eval(t, deriv)
        t = torch.as_tensor(t, dtype=self._derivs.dtype, device=self._derivs.device)
        coeffs = self._coeffs.detach().contiguous()
        n, K, C = _flatten(coeffs)
        tq = t.detach().reshape(-1).contiguous()
        out = torch.empty(*coeffs.shape[:-2], tq.numel(), C, dtype=coeffs.dtype, device=coeffs.device)
        knots = self._t.detach().to(coeffs.dtype).contiguous()
        _capi.check(_capi.lib().ncde_path_eval_smooth(_capi.dtype_code(coeffs), coeffs.data_ptr(), self._derivs.data_ptr(),
                                                      knots.data_ptr(), self.gradient_matching_coeffs.data_ptr(),
                                                      self.match_terms, float(self.gradient_matching_eps), n, K, C,
                                                      tq.data_ptr(), tq.numel(), int(deriv), out.data_ptr(),
                                                      _capi.stream_ptr(coeffs.device)))
        return out.reshape(*coeffs.shape[:-2], *t.shape, C)


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
