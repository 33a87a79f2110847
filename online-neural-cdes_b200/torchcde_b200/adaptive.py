"""dopri5 (adaptive Dormand-Prince) through libncde_b200: every step-control decision is taken on the device
(csrc/adaptive_kernels.cuh); this module only prepares the problem description, wires autograd and checks the status
flags.

Replaces torchdiffeq's RKAdaptiveStepsizeODESolver / Dopri5Solver (modules/torchdiffeq/torchdiffeq/_impl/
rk_common.py:117-313, dopri5.py) for the CDE vector field, and — for adjoint=True — OdeintAdjointMethod with dopri5 as
the adjoint method (adjoint.py:9-215).  Backprop *through* the adaptive solver (adjoint=False with gradients) is not
implemented and raises.
"""
import ctypes
import math
import warnings

import numpy as np
import torch

from . import _capi

_KNOWN = ("min_step", "max_step", "first_step", "safety", "ifactor", "dfactor", "max_num_steps", "dtype", "norm",
          "step_t", "jump_t")

last_stats = {}
last_adjoint_stats = {}


def _check_options(options):
    for k in ("step_t", "jump_t"):
        if options.get(k) is not None:
            raise NotImplementedError("options['{}'] is not implemented".format(k))
    if "norm" in options:
        raise NotImplementedError("a custom error norm is not implemented (the default RMS norm over the batch is)")
    if options.get("dtype", torch.float64) != torch.float64:
        raise NotImplementedError("time scalars are float64 like the reference default")
    unused = {k: v for k, v in options.items() if k not in _KNOWN}
    if unused:
        warnings.warn("Dopri5Solver: Unexpected arguments {}".format(unused))


def _fill_adaptive(problem, rtol, atol, options, T, out_t):
    ad = problem.adaptive
    ad.rtol, ad.atol = float(rtol), float(atol)
    ad.min_step = float(options.get("min_step", 0.0))
    ad.max_step = float(options.get("max_step", math.inf))
    fs = options.get("first_step")
    ad.first_step = -1.0 if fs is None else float(fs)
    ad.safety = float(options.get("safety", 0.9))
    ad.ifactor = float(options.get("ifactor", 10.0))
    ad.dfactor = float(options.get("dfactor", 0.2))
    cap = options.get("max_num_steps")
    ad.max_attempts = int(min(1_000_000, cap * max(T - 1, 1))) if cap is not None else 1_000_000
    ad.n_out = T
    ad.out_t = out_t.ctypes.data_as(ctypes.c_void_p)
    return ad


class _NoGrid:
    n_steps = 0
    stage_t = dt = np.zeros(1, dtype=np.float32)
    out_step = np.zeros(1, dtype=np.int64)
    out_mode = np.zeros(1, dtype=np.int32)
    out_slope = np.zeros(1, dtype=np.float32)

    def __init__(self, T):
        self.n_out = T


def _check_flags(flags, attempted, cap):
    # rk_common.py:196-197, 232-233
    assert not (flags & _capi.FLAG_MAX_STEPS), "max_num_steps exceeded ({}>={})".format(attempted, cap)
    assert not (flags & _capi.FLAG_DT_UNDERFLOW), "underflow in dt"
    assert not (flags & _capi.FLAG_NONFINITE), "non-finite values in state `y`"


def _forward(Xf, spec, z0f, t_host, precision, rtol, atol, options):
    from . import solver as S
    B, H = z0f.shape
    C = spec.weights[-1].shape[0] // H
    T = int(t_host.numel())
    out_t = np.ascontiguousarray(t_host.to(torch.float64).numpy())
    problem, keep = S._build_problem(Xf, spec, B, H, C, "dopri5", S._PRECISIONS[precision], _NoGrid(T))
    ad = _fill_adaptive(problem, rtol, atol, options, T, out_t)
    L = _capi.lib()
    dev = z0f.device
    z_out = torch.empty(T, B, H, dtype=torch.float32, device=dev)
    wbytes = L.ncde_solve_workspace_bytes(ctypes.byref(problem), 0)
    work = torch.empty(max(wbytes, 16), dtype=torch.uint8, device=dev)
    stats = torch.zeros(8 + 192, dtype=torch.int64, device=dev)
    launches = ctypes.c_int64(0)
    _capi.check(L.ncde_solve_adaptive_fwd(ctypes.byref(problem), z0f.data_ptr(), z_out.data_ptr(), work.data_ptr(),
                                          wbytes, stats.data_ptr(), ctypes.byref(launches), _capi.stream_ptr(dev)))
    S.last_launches["fwd"] = launches.value
    host = stats.cpu()   # the one synchronisation of the forward solve
    attempted, accepted, nfe, flags = [int(v) for v in host[:4].tolist()]
    last_stats.clear()
    last_stats.update(attempted=attempted, accepted=accepted, nfe=nfe, flags=flags, launches=launches.value,
                      first_step=float(host[4:5].view(torch.float64)[0]),
                      init_h0_d0_d1_d2=host[5:7].view(torch.float32).tolist(),
                      trace=host[8:8 + 3 * min(attempted, 64)].view(torch.float64).view(-1, 3).tolist())
    _check_flags(flags, attempted, ad.max_attempts)
    return z_out, nfe


class _AdjointAdaptiveSolve(torch.autograd.Function):
    """Forward: device-controlled dopri5, nothing saved but the outputs.  Backward: ncde_solve_adjoint_adaptive_bwd."""

    @staticmethod
    def forward(ctx, Xf, spec, precision, fwd_args, adj_args, t_host, func, z0f, coeffs_for_graph, *params):
        rtol, atol, options = fwd_args
        z_out, nfe = _forward(Xf, spec, z0f.detach().contiguous(), t_host, precision, rtol, atol, options)
        if hasattr(func, "nfe"):
            func.nfe += nfe
        ctx.Xf, ctx.spec, ctx.precision, ctx.adj_args, ctx.t_host, ctx.func = Xf, spec, precision, adj_args, t_host, func
        ctx.save_for_backward(z_out)
        return z_out

    @staticmethod
    def backward(ctx, grad_out):
        from . import solver as S
        (z_out,) = ctx.saved_tensors
        spec, t_host = ctx.spec, ctx.t_host
        if ctx.needs_input_grad[8]:
            raise NotImplementedError("gradients with respect to the control path coefficients are not implemented")
        T, B, H = z_out.shape
        C = spec.weights[-1].shape[0] // H
        dev = grad_out.device
        rtol, atol, options = ctx.adj_args
        out_t = np.ascontiguousarray(t_host.to(torch.float64).numpy())
        problem, keep = S._build_problem(ctx.Xf, spec, B, H, C, "dopri5", S._PRECISIONS[ctx.precision], _NoGrid(T))
        ad = _fill_adaptive(problem, rtol, atol, options, T, out_t)
        L = _capi.lib()
        g = grad_out.contiguous()
        grad_z0 = torch.empty(B, H, dtype=torch.float32, device=dev)
        uniq = spec.unique_params
        grads = [torch.zeros_like(p, dtype=torch.float32, memory_format=torch.contiguous_format) for p, _, _ in uniq]
        by_layer_w, by_layer_b = {}, {}
        for gt, (_, kind, layer) in zip(grads, uniq):
            (by_layer_w if kind == "W" else by_layer_b)[layer] = gt
        gW = (ctypes.c_void_p * _capi.MAX_LAYERS)()
        gb = (ctypes.c_void_p * _capi.MAX_LAYERS)()
        first_of_slot = {}
        for i in range(len(spec.weights)):
            j = first_of_slot.setdefault(spec.slots[i], i)
            gW[i] = by_layer_w[j].data_ptr()
            gb[i] = by_layer_b[j].data_ptr() if j in by_layer_b else None
        wbytes = L.ncde_solve_adjoint_adaptive_workspace_bytes(ctypes.byref(problem))
        work = torch.empty(max(wbytes, 16), dtype=torch.uint8, device=dev)
        stats = torch.zeros(200, dtype=torch.int64, device=dev)
        launches = ctypes.c_int64(0)
        _capi.check(L.ncde_solve_adjoint_adaptive_bwd(ctypes.byref(problem), z_out.data_ptr(), g.data_ptr(),
                                                      grad_z0.data_ptr(), gW, gb, work.data_ptr(), wbytes,
                                                      stats.data_ptr(), ctypes.byref(launches), _capi.stream_ptr(dev)))
        S.last_launches["bwd"] = launches.value
        host = stats.cpu()
        attempted, accepted, nfe, flags = [int(v) for v in host[:4].tolist()]
        last_adjoint_stats.clear()
        last_adjoint_stats.update(attempted=attempted, accepted=accepted, nfe=nfe, flags=flags, launches=launches.value,
                                  trace=host[8:8 + 3 * min(attempted, 64)].view(torch.float64).view(-1, 3).tolist())
        _check_flags(flags, attempted, ad.max_attempts)
        if hasattr(ctx.func, "nfe"):
            ctx.func.nfe += nfe
        return (None, None, None, None, None, None, None, grad_z0, None) + tuple(grads)


def solve(X, func, spec, z0, t, t_host, adjoint, options, kwargs, coeffs=None):
    """``coeffs``: the coefficient tensor cdeint prepared for the autograd graph (detached when the path is not listed in
    adjoint_params, solver.py:201-221); defaults to X._coeffs."""
    from . import solver as S
    options = dict(options)
    precision = options.pop("precision", S.default_precision)
    _check_options(options)
    if coeffs is None:
        coeffs = X._coeffs
    needs_grad = torch.is_grad_enabled() and (z0.requires_grad or coeffs.requires_grad or
                                              any(p.requires_grad for p, _, _ in spec.unique_params))
    batch_shape = z0.shape[:-1]
    H = z0.shape[-1]
    Xf = X if len(batch_shape) == 1 else S._flatten_path(X)
    T = int(t_host.numel())
    z0f = z0.reshape(-1, H)
    if not needs_grad:
        z_out, nfe = _forward(Xf, spec, z0f.detach().contiguous(), t_host, precision, kwargs["rtol"], kwargs["atol"], options)
        if hasattr(func, "nfe"):
            func.nfe += nfe
        return z_out.reshape(T, *batch_shape, H)
    if not adjoint:
        raise NotImplementedError("backprop through the adaptive solver (method='dopri5', adjoint=False) is not "
                                  "implemented; use adjoint=True (continuous adjoint) or method='rk4'")
    # adjoint.py:159-171
    adj_method = kwargs.get("adjoint_method") or "dopri5"
    if adj_method != "dopri5":
        raise NotImplementedError("adjoint_method '{}' after a dopri5 forward solve is not implemented".format(adj_method))
    adj_rtol = kwargs.get("adjoint_rtol")
    adj_atol = kwargs.get("adjoint_atol")
    adj_rtol = kwargs["rtol"] if adj_rtol is None else adj_rtol
    adj_atol = kwargs["atol"] if adj_atol is None else adj_atol
    adj_options = kwargs.get("adjoint_options")
    if adj_options is None:
        adj_options = {k: v for k, v in options.items() if k != "norm"}
    else:
        adj_options = dict(adj_options)
        adj_options.pop("norm", None) if adj_options.get("norm") is None else None
        _check_options(adj_options)
    params = [p for p, _, _ in spec.unique_params]
    out = _AdjointAdaptiveSolve.apply(Xf, spec, precision, (kwargs["rtol"], kwargs["atol"], options),
                                      (adj_rtol, adj_atol, adj_options), t_host, func, z0f, coeffs, *params)
    return out.reshape(T, *batch_shape, H)
