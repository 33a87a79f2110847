"""Linear / rectilinear interpolation: coefficient construction and evaluation on the GPU.

Mirrors torchcde/interpolation_linear.py of the reference (names, arguments, return conventions, errors).  All
arithmetic runs in libncde_b200 (csrc/interp.cu); results are bit-identical to the reference on the same inputs.
"""
import warnings

import torch

from . import _capi
from . import interpolation_base
from . import misc


def _flatten(x):
    L, C = x.size(-2), x.size(-1)
    n = x.numel() // (L * C) if x.numel() else 0
    return n, L, C


def _prepare_rectilinear_interpolation(data, time_index):
    """(…, L, C) -> (…, 2L-1, C).  torchcde/interpolation_linear.py:87-128."""
    n_channels = data.size(-1)
    assert isinstance(time_index, int), "Index of the time channel must be an integer in [0, {}]".format(n_channels - 1)
    assert 0 <= time_index < n_channels, "Time index must be in [0, {}], was given {}.".format(n_channels - 1,
                                                                                               time_index)
    _capi.require_cuda(data)
    x = data.contiguous()
    n, L, C = _flatten(x)
    out = torch.empty(*x.shape[:-2], 2 * L - 1, C, dtype=x.dtype, device=x.device)
    flags = torch.zeros(1, dtype=torch.int32, device=x.device)
    _capi.check(_capi.lib().ncde_rectilinear_prepare(_capi.dtype_code(x), x.data_ptr(), out.data_ptr(), n, L, C,
                                                     time_index, flags.data_ptr(), _capi.stream_ptr(x.device)))
    # the reference asserts on NaN times before doing any work; reading the flag is the one sync of this constructor
    assert not (int(flags.item()) & _capi.FLAG_NAN_TIME), \
        "There exist nan values in the time column which is not allowed. If the times are padded with nans after " \
        "final time, a simple solution is to forward fill the final time."
    return out


def linear_interpolation_coeffs(x, t=None, rectilinear=None, initial_value_if_nan=None, forward_fill=False):
    """Knots of the linear (or rectilinear) interpolation of a batch of paths; NaN = missing value.

    Same contract as torchcde/interpolation_linear.py:131-180: returns the input object itself when there is
    nothing to do, mutates ``x`` in place when ``initial_value_if_nan`` is given, warns when a rectilinear path
    starts with missing values, raises ValueError for malformed input and AssertionError for NaN times.
    """
    _capi.require_cuda(x)
    if initial_value_if_nan is not None:
        x[..., 0, :][torch.isnan(x[..., 0, :])] = initial_value_if_nan

    if rectilinear is not None:
        if torch.isnan(x[..., 0, :]).any():
            warnings.warn("The data `x` begins with missing values in some channels. The path will be constructed by "
                          "backward-filling the first observed value, which is not causal. Raising a warning as the "
                          "`rectilinear` argument has also been passed, which is nearly always only used when "
                          "causality is desired. If you need causality then fill in the missing value at the start of "
                          "each channel with whatever you'd like it to be. (The mean over that channel is a common "
                          "choice.)")
        x = _prepare_rectilinear_interpolation(x, rectilinear)

    if forward_fill:
        x = misc.forward_fill(x)

    t = misc.validate_input_path(x, t)

    if torch.isnan(x).any():
        # the reference leaves the caller's tensor untouched on this path
        x = x.clone(memory_format=torch.contiguous_format) if not (rectilinear is not None or forward_fill) \
            else x.contiguous()
        n, L, C = _flatten(x)
        tt = t.to(device=x.device, dtype=x.dtype).contiguous()
        _capi.check(_capi.lib().ncde_linear_fill_missing(_capi.dtype_code(x), x.data_ptr(), tt.data_ptr(), n, L, C,
                                                         _capi.stream_ptr(x.device)))
    return x


class LinearInterpolation(interpolation_base.InterpolationBase):
    """Linear interpolation of a batch of controls and its derivative (torchcde/interpolation_linear.py:183-234)."""

    def __init__(self, coeffs, t=None, **kwargs):
        super(LinearInterpolation, self).__init__(**kwargs)
        _capi.require_cuda(coeffs)
        if t is None:
            t = misc.default_times(coeffs.size(-2), coeffs.dtype, coeffs.device)
        t_dev = t.to(coeffs.device)
        if t_dev is not t:
            misc.attach_host(t_dev, misc.host_values(t))
        coeffs_c = coeffs.detach().contiguous()
        n, K, C = _flatten(coeffs_c)
        derivs = torch.empty(*coeffs_c.shape[:-2], K - 1, C, dtype=coeffs_c.dtype, device=coeffs_c.device)
        tt = t_dev.detach().to(coeffs_c.dtype).contiguous()
        _capi.check(_capi.lib().ncde_linear_derivs(_capi.dtype_code(coeffs_c), coeffs_c.data_ptr(), tt.data_ptr(),
                                                   derivs.data_ptr(), n, K, C, _capi.stream_ptr(coeffs_c.device)))
        misc.host_values(t_dev)   # mirror the knots on the host once (keyed on storage + version)
        self.register_buffer('_t', t_dev)
        self.register_buffer('_coeffs', coeffs)
        self.register_buffer('_derivs', derivs)

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

    def _eval(self, t, deriv):
        t = torch.as_tensor(t, dtype=self._derivs.dtype, device=self._derivs.device)
        coeffs = self._coeffs.detach().contiguous()
        C = coeffs.size(-1)
        tq = t.detach().reshape(-1).contiguous()
        knots = self._t.detach().to(coeffs.dtype).contiguous()
        if self._coeffs.requires_grad and torch.is_grad_enabled():
            out = interpolation_base.PathEvalGrad.apply(self._coeffs, _capi.PATH_LINEAR, self._derivs, knots, tq,
                                                        bool(deriv), C)
        else:
            out = interpolation_base.path_eval_raw(_capi.PATH_LINEAR, coeffs, self._derivs, knots, tq, deriv, C)
        return out.reshape(*coeffs.shape[:-2], *t.shape, C)

    def evaluate(self, t):
        return self._eval(t, False)

    def derivative(self, t):
        return self._eval(t, True)

    def knot_index(self, t):
        """bucketize(t, knots) - 1 clamped to [0, K-2] (interpolation_linear.py:212-219), from the device kernel."""
        t = torch.as_tensor(t, dtype=self._derivs.dtype, device=self._derivs.device)
        tq = t.reshape(-1).contiguous()
        coeffs = self._coeffs.detach().contiguous()
        n, K, C = _flatten(coeffs)
        out = torch.empty(*coeffs.shape[:-2], tq.numel(), C, dtype=coeffs.dtype, device=coeffs.device)
        index = torch.empty(tq.numel(), dtype=torch.int64, device=coeffs.device)
        knots = self._t.detach().to(coeffs.dtype).contiguous()
        _capi.check(_capi.lib().ncde_path_eval(_capi.PATH_LINEAR, _capi.dtype_code(coeffs), coeffs.data_ptr(),
                                               self._derivs.data_ptr(), knots.data_ptr(), n, K, C, tq.data_ptr(),
                                               tq.numel(), 1, out.data_ptr(), index.data_ptr(),
                                               _capi.stream_ptr(coeffs.device)))
        return index.reshape(t.shape)
