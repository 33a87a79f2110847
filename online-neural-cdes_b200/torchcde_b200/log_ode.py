"""Log-ODE transform: log-signatures over windows (depth 1, 2 and 3) on the GPU.

Mirrors torchcde/log_ode.py of the reference (`logsig_windows`, the deprecated `logsignature_windows`).  The reference
delegates the log-signature itself to the third-party `signatory` extension (un-vendored, un-pinned: SURVEY §8c); here
depths 1 to 3 are computed by `ncde_logsig_windows` (csrc/interp.cu) in Signatory's default "words" channel order — the
coefficients of log(signature) at the Lyndon words of length 1 (increments), 2 ((i, j), i < j: the Levy areas) and 3
((i, j, k), j >= i, k > i).  Deeper log-signatures raise NotImplementedError.
The window bookkeeping follows log_ode.py:15-47 line by line (it decides which rows exist, so it must not drift);
missing values and the inserted window boundaries are filled through `linear_interpolation_coeffs` as in :47.
"""
import torch

from . import _capi
from . import interpolation_linear
from . import misc


def logsignature_channels(in_channels, depth):
    """signatory.logsignature_channels for depth <= 3 (Witt's formula: d, d (d - 1) / 2, (d^3 - d) / 3 Lyndon words)."""
    d = in_channels
    if depth not in (1, 2, 3):
        raise NotImplementedError("log-signatures of depth {} are not implemented (depth 1, 2 and 3 are)".format(depth))
    return d + (d * (d - 1) // 2 if depth >= 2 else 0) + ((d ** 3 - d) // 3 if depth >= 3 else 0)


def _logsignature_windows(x, depth, window_length, t, _version):
    _capi.require_cuda(x)
    logsignature_channels(x.size(-1), depth)   # depth check before any work
    t = misc.validate_input_path(x, t)

    # log_ode.py:18-23
    timespan = t[-1] - t[0]
    num_pieces = (timespan / window_length).ceil().to(int).item()
    end_t = t[0] + num_pieces * window_length
    new_t = torch.linspace(t[0], end_t, num_pieces + 1, dtype=t.dtype, device=t.device)
    new_t = torch.min(new_t, t.max())

    # log_ode.py:25-38 — on host copies of the (short) time vectors
    t_host, new_t_host = t.detach().cpu(), new_t.detach().cpu()
    t_index = 0
    new_t_unique = []
    new_t_indices = []
    for new_t_elem in new_t_host:
        while True:
            lequal = bool(new_t_elem <= t_host[t_index])
            close = bool(new_t_elem.allclose(t_host[t_index]))
            if lequal or close:
                break
            t_index += 1
        new_t_indices.append(t_index + len(new_t_unique))
        if close:
            continue
        new_t_unique.append(new_t_elem.unsqueeze(0))

    batch_dimensions = x.shape[:-2]
    missing_X = torch.full((1,), float('nan'), dtype=x.dtype, device=x.device).expand(*batch_dimensions, 1, x.size(-1))
    if len(new_t_unique) > 0:
        t, indices = torch.cat([t, *[e.to(t.device) for e in new_t_unique]]).sort()
        x = torch.cat([x, missing_X], dim=-2)[..., indices.clamp(0, x.size(-2)), :]

    # log_ode.py:47
    x = interpolation_linear.linear_interpolation_coeffs(x, t).contiguous()

    d = x.size(-1)
    ch = logsignature_channels(d, depth)
    W = len(new_t_indices) - 1
    n = x.numel() // (x.size(-2) * d) if x.numel() else 0
    out = torch.empty(*batch_dimensions, W + 1, ch, dtype=x.dtype, device=x.device)
    idx = torch.tensor(new_t_indices, dtype=torch.int32, device=x.device)
    wscale = None
    if _version == 0:
        wscale = (new_t[1:] - new_t[:-1]).to(x.dtype).contiguous()
    elif _version != 1:
        raise RuntimeError
    _capi.check(_capi.lib().ncde_logsig_windows(_capi.dtype_code(x), x.data_ptr(), idx.data_ptr(),
                                                None if wscale is None else wscale.data_ptr(), out.data_ptr(), n,
                                                x.size(-2), d, depth, W, _capi.stream_ptr(x.device)))
    if _version == 0:
        return out, new_t
    return out


def logsignature_windows(x, depth, window_length, t=None):
    """DEPRECATED variant kept for backward compatibility (log_ode.py:80-109): every window's log-signature is scaled by
    the window duration; returns (values, times)."""
    return _logsignature_windows(x, depth, window_length, t, _version=0)


def logsig_windows(x, depth, window_length, t=None):
    """Log-signatures over windows of length `window_length`, cumulatively summed, as in the log-ODE method
    (log_ode.py:112-136).  x: (..., length, channels), NaN = missing.  Returns (..., windows + 1, logsig channels); the
    corresponding times are 0, 1, ..., windows."""
    return _logsignature_windows(x, depth, window_length, t, _version=1)
