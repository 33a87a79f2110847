"""Natural cubic splines: coefficient construction (batched Thomas solve) and evaluation on the GPU.

Mirrors torchcde/interpolation_cubic.py of the reference.
"""
import torch

from . import _capi
from . import interpolation_base
from . import misc


def _natural_cubic_spline_coeffs(x, t, _version):
    _capi.require_cuda(x)
    t = misc.validate_input_path(x, t)
    xc = x.detach().contiguous()
    L, C = xc.size(-2), xc.size(-1)
    n = xc.numel() // (L * C) if xc.numel() else 0
    out = torch.empty(*xc.shape[:-2], L - 1, 4 * C, dtype=xc.dtype, device=xc.device)
    tt = t.detach().to(device=xc.device, dtype=xc.dtype).contiguous()
    code = _capi.dtype_code(xc)
    scratch = torch.empty(_capi.lib().ncde_cubic_scratch_bytes(code, n, L, C), dtype=torch.uint8, device=xc.device)
    _capi.check(_capi.lib().ncde_natural_cubic_coeffs(code, xc.data_ptr(), tt.data_ptr(), out.data_ptr(), n, L, C,
                                                      _version, scratch.data_ptr(), _capi.stream_ptr(xc.device)))
    return out


def natural_cubic_spline_coeffs(x, t=None):
    """DEPRECATED variant kept for compatibility (torchcde/interpolation_cubic.py:192-230): missing values at the
    two ends are imputed only at the very first / last time."""
    return _natural_cubic_spline_coeffs(x, t, _version=0)


def natural_cubic_coeffs(x, t=None):
    """Coefficients of the natural cubic spline through a batch of paths, shape (..., length - 1, 4 * channels)
    = [a | b | 2c | 3d] (torchcde/interpolation_cubic.py:233-265).  Missing values (NaN) are supported; leading and
    trailing gaps are filled from the first / last observation."""
    return _natural_cubic_spline_coeffs(x, t, _version=1)


class NaturalCubicSpline(interpolation_base.InterpolationBase):
    """Natural cubic spline and its derivative (torchcde/interpolation_cubic.py:268-336)."""

    def __init__(self, coeffs, t=None, **kwargs):
        super(NaturalCubicSpline, self).__init__(**kwargs)
        _capi.require_cuda(coeffs)
        if t is None:
            t = misc.default_times(coeffs.size(-2) + 1, coeffs.dtype, coeffs.device)
        t_dev = t.to(coeffs.device)
        if t_dev is not t:
            misc.attach_host(t_dev, misc.host_values(t))
        channels = coeffs.size(-1) // 4
        if channels * 4 != coeffs.size(-1):
            raise ValueError("Passed invalid coeffs.")
        misc.host_values(t_dev)   # mirror the knots on the host once (keyed on storage + version)
        self._channels = channels
        self.register_buffer('_t', t_dev)
        self.register_buffer('_coeffs', coeffs)

    # the reference registers the four slices as buffers; expose them as views of the packed tensor
    @property
    def _a(self):
        return self._coeffs[..., :self._channels]

    @property
    def _b(self):
        return self._coeffs[..., self._channels:2 * self._channels]

    @property
    def _two_c(self):
        return self._coeffs[..., 2 * self._channels:3 * self._channels]

    @property
    def _three_d(self):
        return self._coeffs[..., 3 * self._channels:]

    @property
    def _t_host(self):
        """Host mirror of the knots; refreshed when the buffer was edited in place or reloaded."""
        return misc.host_values(self._t)

    @property
    def grid_points(self):
        return misc.attach_host(self._t, self._t_host)

    @property
    def interval(self):
        host = torch.stack([self._t_host[0], self._t_host[-1]])
        return misc.attach_host(host.to(self._t.device), host)

    def _eval(self, t, deriv, want_index=False):
        t = torch.as_tensor(t, dtype=self._coeffs.dtype, device=self._coeffs.device)
        coeffs = self._coeffs.detach().contiguous()
        C = self._channels
        tq = t.detach().reshape(-1).contiguous()
        knots = self._t.detach().to(coeffs.dtype).contiguous()
        if want_index:
            index = torch.empty(tq.numel(), dtype=torch.int64, device=coeffs.device)
            interpolation_base.path_eval_raw(_capi.PATH_CUBIC, coeffs, None, knots, tq, deriv, C, index)
            return index.reshape(t.shape)
        if self._coeffs.requires_grad and torch.is_grad_enabled():
            out = interpolation_base.PathEvalGrad.apply(self._coeffs, _capi.PATH_CUBIC, None, knots, tq, bool(deriv), C)
        else:
            out = interpolation_base.path_eval_raw(_capi.PATH_CUBIC, coeffs, None, knots, tq, deriv, C)
        return out.reshape(*coeffs.shape[:-2], *t.shape, C)

    def evaluate(self, t):
        return self._eval(t, False)

    def derivative(self, t):
        return self._eval(t, True)

    def knot_index(self, t):
        return self._eval(t, True, want_index=True)


# experiments/sim_bm_toy_example.py:46 calls torchcde.CubicSpline, which the vendored 0.2.0 does not export
CubicSpline = NaturalCubicSpline
