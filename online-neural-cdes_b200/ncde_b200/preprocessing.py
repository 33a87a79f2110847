"""Offline preprocessing and loader helpers over RAGGED series, mirroring

* ``get_data/transformers.py:7-85``  ``Interpolation`` (linear / rectilinear / cubic / linear_forward_fill coefficients of a list of
  series of different lengths — a Python loop over the series in the reference, "likely to take a LONG time",
  get_data/mimic-iv/prepare.py:262),
* ``get_data/common.py:59-80``  ``temporal_pipeline``,
* ``experiments/ingredients/loader.py:100-113``  the ``rectilinear-intensity`` channels,
* ``experiments/ingredients/loader.py:158-166,181-202``  sorting by length and per-batch padding (PadRaggedTensors + ForwardFill).

All series are packed into one NaN-padded (n, Lmax, C) device tensor with a length per series and go through ONE call of
``ncde_ragged_interpolate`` (csrc/interp.cu); results are bit-identical to the per-series reference calls.  The host side only
packs / slices memory.  No CPU fallback: the library must be built and a CUDA device present.
"""
import warnings

import numpy as np
import torch

from torchcde_b200 import _capi

_METHODS = {"linear": _capi.RAGGED_LINEAR, "linear_forward_fill": _capi.RAGGED_LINEAR,
            "rectilinear": _capi.RAGGED_RECTILINEAR, "cubic": _capi.RAGGED_CUBIC}


def _out_rows(method, L):
    return {_capi.RAGGED_LINEAR: L, _capi.RAGGED_RECTILINEAR: 2 * L - 1, _capi.RAGGED_CUBIC: L - 1}[method]


def ragged_interpolate(data, method="linear", initial_nan_to_zero=True, intensity=False, pad=False, device=None):
    """Coefficients of every series of ``data`` (a list of (L_i, C) tensors or one (n, L, C) tensor) in one launch.

    Returns ``(coeffs, rows)``: ``coeffs`` (n, Kmax, Cout) on the GPU, series i occupying its first ``rows[i]`` rows
    (the rest repeats its last row if ``pad`` else is NaN); Cout = C (linear), C or 2C-1 with ``intensity`` (rectilinear),
    4C (cubic).
    """
    code = _METHODS[method]
    is_list = not isinstance(data, torch.Tensor)
    series = list(data) if is_list else None
    if is_list:
        if not series:
            raise ValueError("no series given")
        lengths = [int(d.shape[0]) for d in series]
        C = int(series[0].shape[-1])
        for d in series:
            if d.dim() != 2 or d.shape[-1] != C:
                raise ValueError("every series must have shape (length, {})".format(C))
            if not d.is_floating_point():
                raise ValueError("The input must be a floating point tensor.")
        x = torch.nn.utils.rnn.pad_sequence(series, batch_first=True, padding_value=float("nan"))
    else:
        if data.dim() != 3:
            raise ValueError("a tensor input must have shape (n, length, channels)")
        if not data.is_floating_point():
            raise ValueError("The input must be a floating point tensor.")
        x = data
        lengths = [int(data.shape[1])] * int(data.shape[0])
        C = int(data.shape[2])
    if min(lengths) < 2:
        raise ValueError("Must have a time dimension of size at least 2.")
    if device is None:
        device = x.device if x.is_cuda else torch.device("cuda")
    x = x.to(device).contiguous()
    _capi.require_cuda(x)
    n, Lmax = x.shape[0], x.shape[1]
    if code == _capi.RAGGED_RECTILINEAR and not initial_nan_to_zero and bool(torch.isnan(x[:, 0]).any()):
        warnings.warn("The data `x` begins with missing values in some channels. The path will be constructed by "
                      "backward-filling the first observed value, which is not causal.")
    len_dev = torch.tensor(lengths, dtype=torch.int32, device=device)
    Cout = (2 * C - 1 if intensity else C) if code == _capi.RAGGED_RECTILINEAR else (4 * C if code == _capi.RAGGED_CUBIC else C)
    out = torch.empty(n, _out_rows(code, Lmax), Cout, dtype=x.dtype, device=device)
    L_ = _capi.lib()
    dt = _capi.dtype_code(x)
    scratch = torch.empty(L_.ncde_ragged_scratch_bytes(code, dt, n, Lmax, C), dtype=torch.uint8, device=device)
    flags = torch.zeros(1, dtype=torch.int32, device=device)
    _capi.check(L_.ncde_ragged_interpolate(code, dt, x.data_ptr(), len_dev.data_ptr(), out.data_ptr(), n, Lmax, C, 0,
                                           int(bool(initial_nan_to_zero)), int(bool(intensity)), int(bool(pad)),
                                           scratch.data_ptr(), flags.data_ptr(), _capi.stream_ptr(device)))
    if code == _capi.RAGGED_RECTILINEAR:
        assert not (int(flags.item()) & _capi.FLAG_NAN_TIME), \
            "There exist nan values in the time column which is not allowed. If the times are padded with nans after " \
            "final time, a simple solution is to forward fill the final time."
    return out, [_out_rows(code, L) for L in lengths]


class Interpolation:
    """Linear, rectilinear, cubic schemes over a list of series — same constructor arguments, ``fit`` / ``transform`` /
    ``fit_transform`` contract and in-place "causality" side effect as get_data/transformers.py:7-85."""

    def __init__(self, method="linear", channel_indices=None, initial_nan_to_zero=True, return_as_list=True):
        assert method in ["linear", "rectilinear", "cubic", "hybrid", "linear_forward_fill"], \
            "Got method {} which is not recognised".format(method)
        if method == "hybrid":
            assert channel_indices is not None, "Hybrid requires specification of the hybrid indices."
            raise NotImplementedError
        self.method = method
        self.channel_indices = channel_indices
        self.initial_nan_to_zero = initial_nan_to_zero
        self.return_as_list = return_as_list
        self._rectilinear = 0 if self.method == "rectilinear" else None

    def __repr__(self):
        return "{} Interpolation".format(self.method.title())

    def fit(self, data, labels=None):
        return self

    def fit_transform(self, data, labels=None):
        return self.fit(data, labels).transform(data)

    def transform(self, data):
        # transformers.py:52-55 — the reference zeroes the callers' first rows in place (its `temporal_data_raw` and the loader's
        # intensity channels rely on that), so the side effect is kept; the kernel applies the same rule to its own copy
        if self.initial_nan_to_zero:
            for d in data:
                d[:1, :][torch.isnan(d[:1, :])] = 0.0
        is_tensor = isinstance(data, torch.Tensor)
        src_device = data.device if is_tensor else data[0].device
        coeffs, rows = ragged_interpolate(data, self.method, self.initial_nan_to_zero)
        coeffs = coeffs.to(src_device)
        if is_tensor:
            return coeffs
        return [coeffs[i, :k] for i, k in enumerate(rows)]


def temporal_pipeline(temporal_data, interpolation_method="linear", return_as_numpy=True):
    """get_data/common.py:59-80."""
    assert len(temporal_data[0].shape) == 2
    temporal_out = Interpolation(method=interpolation_method).fit_transform(temporal_data)
    if return_as_numpy:
        if all([len(x) == len(temporal_out[0]) for x in temporal_out]):
            temporal_out = np.stack([x.cpu().numpy() for x in temporal_out]).astype(np.float32)
        else:
            temporal_out = [x.cpu().numpy().astype(np.float32) for x in temporal_out]
    return temporal_out


def rectilinear_intensity(raw_data, initial_nan_to_zero=True, pad=False):
    """Rectilinear coefficients with the observation-intensity channels appended, from the RAW series — what
    experiments/ingredients/loader.py:100-113 assembles per series from `temporal_data_rectilinear` and `temporal_data_raw`.
    Returns ``(coeffs (n, 2Lmax-1, 2C-1) on the GPU, rows)``."""
    return ragged_interpolate(raw_data, "rectilinear", initial_nan_to_zero, intensity=True, pad=pad)


def sort_unequal_lengths(static, temporal, labels):
    """experiments/ingredients/loader.py:158-166: shortest series first."""
    lengths = [len(x) for x in temporal]
    idx = sorted(range(len(lengths)), key=lambda k: lengths[k])
    static = static[idx] if static is not None else None
    temporal = [temporal[i] for i in idx] if isinstance(temporal, list) else temporal[idx]
    labels = [labels[i] for i in idx] if isinstance(labels, list) else labels[idx]
    return static, temporal, labels, idx


def padded_batches(coeffs, rows, batch_size):
    """Batches of a length-sorted coefficient set: batch j = series [j*batch_size, (j+1)*batch_size) cut to the longest of them,
    shorter series repeating their last row — PadRaggedTensors + ForwardFill per batch (loader.py:181-202).  ``coeffs`` must
    come from ``ragged_interpolate(..., pad=True)``; the batches are views, nothing is copied."""
    assert all(rows[i] <= rows[i + 1] for i in range(len(rows) - 1)), \
        "Data is of unequal length and has not been sorted. This will lead to slow training, please sort the data in " \
        "order of length first."
    return [coeffs[i:i + batch_size, :max(rows[i:i + batch_size])] for i in range(0, len(rows), batch_size)]
