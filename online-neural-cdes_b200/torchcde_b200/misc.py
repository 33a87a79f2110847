"""Helpers shared by the interpolation constructors (mirrors torchcde/misc.py of the reference)."""
import torch

from . import _capi


def cheap_stack(tensors, dim):
    """torchcde/misc.py:6-10."""
    if len(tensors) == 1:
        return tensors[0].unsqueeze(dim)
    return torch.stack(tensors, dim=dim)


def host_values(t):
    """Host copy of a small 1-D time tensor.  Tensors created by this package carry a host mirror so the hot path
    never has to synchronise the stream to read them.  The mirror is keyed on the tensor's storage and version
    counter: an in-place edit (or a load_state_dict into a path module) invalidates it and costs one D2H copy."""
    h = getattr(t, "_ncde_host", None)
    key = (t.data_ptr(), t._version)
    if h is None or getattr(t, "_ncde_host_key", key) != key:
        h = t.detach().cpu()
        if h is not t:   # a CPU tensor is its own mirror: nothing to cache
            t._ncde_host, t._ncde_host_key = h, key
    return h


def attach_host(t, host):
    t._ncde_host = host
    t._ncde_host_key = (t.data_ptr(), t._version)
    return t


def default_times(length, dtype, device):
    """t = linspace(0, length-1, length) (torchcde/misc.py:78-79), built on the host and mirrored."""
    host = torch.linspace(0, length - 1, length, dtype=dtype)
    return attach_host(host.to(device), host)


def validate_input_path(x, t):
    """Same checks and ValueErrors as torchcde/misc.py:70-100.  The monotonicity walk runs on a host copy of t."""
    if not x.is_floating_point():
        raise ValueError("X must both be floating point.")
    if x.ndimension() < 2:
        raise ValueError("X must have at least two dimensions, corresponding to time and channels. It instead has "
                         "shape {}.".format(tuple(x.shape)))
    if t is None:
        t = default_times(x.size(-2), x.dtype, x.device)
    if not t.is_floating_point():
        raise ValueError("t must both be floating point.")
    if len(t.shape) != 1:
        raise ValueError("t must be one dimensional. It instead has shape {}.".format(tuple(t.shape)))
    th = host_values(t)
    if th.numel() > 1 and not bool((th[1:] > th[:-1]).all()):
        raise ValueError("t must be monotonically increasing.")
    if x.size(-2) != t.size(0):
        raise ValueError("The time dimension of X must equal the length of t. X has shape {} and t has shape {}, "
                         "corresponding to time dimensions of {} and {} respectively."
                         .format(tuple(x.shape), tuple(t.shape), x.size(-2), t.size(0)))
    if t.size(0) < 2:
        raise ValueError("Must have a time dimension of size at least 2. It instead has shape {}, corresponding to a "
                         "time dimension of size {}.".format(tuple(t.shape), t.size(0)))
    return t


def forward_fill(x, fill_index=-2):
    """Forward fill along ``fill_index`` (torchcde/misc.py:103-126) as one CUDA kernel."""
    assert isinstance(x, torch.Tensor)
    assert x.dim() >= 2 or fill_index == -1
    _capi.require_cuda(x)
    if fill_index < 0:
        fill_index += x.dim()
    if fill_index == x.dim() - 1:
        # fill along the last axis: a (n, L) problem with one channel
        xc = x.contiguous()
        out = torch.empty_like(xc)
        n = xc.numel() // xc.size(-1) if xc.numel() else 0
        _capi.check(_capi.lib().ncde_forward_fill(_capi.dtype_code(xc), xc.data_ptr(), out.data_ptr(), n,
                                                  xc.size(-1), 1, _capi.stream_ptr(x.device)))
        return out
    if fill_index != x.dim() - 2:
        moved = x.movedim(fill_index, -2)
        return forward_fill(moved, -2).movedim(-2, fill_index)
    xc = x.contiguous()
    out = torch.empty_like(xc)
    L, C = xc.size(-2), xc.size(-1)
    n = xc.numel() // (L * C) if xc.numel() else 0
    _capi.check(_capi.lib().ncde_forward_fill(_capi.dtype_code(xc), xc.data_ptr(), out.data_ptr(), n, L, C,
                                              _capi.stream_ptr(x.device)))
    return out


class TupleControl(torch.nn.Module):
    """Several controls over one interval presented as a single control whose evaluate / derivative return tuples
    (the API of torchcde/misc.py:129-166).  The reference fork's cdeint never takes the tuple branch
    (solver.py:198-199 fixes is_tensor=True) and neither does this package's; the class exists so code that builds one
    still imports."""

    def __init__(self, *controls):
        super().__init__()
        if not controls:
            raise ValueError("Expected one or more controls to batch together.")
        first = controls[0]
        for other in controls[1:]:
            if not torch.equal(torch.as_tensor(other.interval), torch.as_tensor(first.interval)):
                raise ValueError("Can only batch together controls over the same interval.")
        shared = all(o.grid_points.shape == first.grid_points.shape and torch.equal(o.grid_points, first.grid_points)
                     for o in controls[1:])
        self._interval = first.interval
        self._grid_points = first.grid_points if shared else None
        self.controls = torch.nn.ModuleList(controls)

    @property
    def interval(self):
        return self._interval

    @property
    def grid_points(self):
        if self._grid_points is None:
            raise RuntimeError("Batch of controls have different grid points.")
        return self._grid_points

    def evaluate(self, t):
        return tuple(c.evaluate(t) for c in self.controls)

    def derivative(self, t):
        return tuple(c.derivative(t) for c in self.controls)
