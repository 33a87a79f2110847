"""dopri5 (adaptive Dormand-Prince) through libncde_b200: every step-control decision is taken on the device
(csrc/adaptive_kernels.cuh); this module only prepares the problem description and checks the status flags.

Replaces torchdiffeq's RKAdaptiveStepsizeODESolver / Dopri5Solver (modules/torchdiffeq/torchdiffeq/_impl/
rk_common.py:117-313, dopri5.py) for the CDE vector field.  Gradients of the adaptive solve (backprop through the
accepted steps, or the continuous adjoint of adjoint.py) are not implemented in this revision and raise.
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


def solve(X, func, spec, z0, t, t_host, adjoint, options, kwargs):
    from . import solver as S
    needs_grad = torch.is_grad_enabled() and (z0.requires_grad or any(w.requires_grad for w in spec.weights))
    if needs_grad:
        raise NotImplementedError("gradients through method='dopri5' are not implemented yet (forward only); "
                                  "wrap the call in torch.no_grad() or use method='rk4'")
    options = dict(options)
    precision = options.pop("precision", S.default_precision)
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

    batch_shape = z0.shape[:-1]
    H = z0.shape[-1]
    z0f = z0.detach().reshape(-1, H).contiguous()
    B = z0f.shape[0]
    C = spec.weights[-1].shape[0] // H
    Xf = X if len(batch_shape) == 1 else S._flatten_path(X)
    T = int(t_host.numel())
    out_t = np.ascontiguousarray(t_host.to(torch.float64).numpy())

    class _NoGrid:
        n_steps, n_out = 0, T
        stage_t = dt = np.zeros(1, dtype=np.float32)
        out_step = np.zeros(1, dtype=np.int64)
        out_mode = np.zeros(1, dtype=np.int32)
        out_slope = np.zeros(1, dtype=np.float32)
    problem, keep = S._build_problem(Xf, spec, B, H, C, "dopri5", S._PRECISIONS[precision], _NoGrid)
    ad = problem.adaptive
    ad.rtol, ad.atol = float(kwargs["rtol"]), float(kwargs["atol"])
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

    L = _capi.lib()
    dev = z0.device
    z_out = torch.empty(T, B, H, dtype=torch.float32, device=dev)
    wbytes = L.ncde_solve_workspace_bytes(ctypes.byref(problem), 0)
    work = torch.empty(max(wbytes, 16), dtype=torch.uint8, device=dev)
    stats = torch.zeros(8 + 192, dtype=torch.int64, device=dev)
    launches = ctypes.c_int64(0)
    _capi.check(L.ncde_solve_adaptive_fwd(ctypes.byref(problem), z0f.data_ptr(), z_out.data_ptr(), work.data_ptr(),
                                          wbytes, stats.data_ptr(), ctypes.byref(launches), _capi.stream_ptr(dev)))
    S.last_launches["fwd"] = launches.value
    host = stats.cpu()   # the one synchronisation of the solve
    attempted, accepted, nfe, flags = [int(v) for v in host[:4].tolist()]
    last_stats.update(first_step=float(host[4:5].view(torch.float64)[0]),
                      init_h0_d0_d1_d2=host[5:7].view(torch.float32).tolist(),
                      trace=host[8:8 + 3 * min(attempted, 64)].view(torch.float64).view(-1, 3).tolist())
    last_stats.update(attempted=attempted, accepted=accepted, nfe=nfe, flags=flags, launches=launches.value)
    # rk_common.py:196-197, 232-233
    assert not (flags & _capi.FLAG_MAX_STEPS), "max_num_steps exceeded ({}>={})".format(attempted, ad.max_attempts)
    assert not (flags & _capi.FLAG_DT_UNDERFLOW), "underflow in dt"
    assert not (flags & _capi.FLAG_NONFINITE), "non-finite values in state `y`"
    if hasattr(func, "nfe"):
        func.nfe += nfe
    return z_out.reshape(T, *batch_shape, H)
