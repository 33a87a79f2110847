"""``cdeint``: drop-in for ``torchcde.cdeint`` (reference: modules/torchcde/torchcde/solver.py:140-238) whose every
integration step runs inside libncde_b200 (hand-written sm_100a CUDA).

The host side only (a) lowers ``func`` to an MLP descriptor, (b) builds the time grid exactly as
``torchdiffeq`` does (modules/torchdiffeq/torchdiffeq/_impl/solvers.py:77-119) and (c) wires the result into
autograd.  There is no eager / CPU fallback: unsupported inputs raise.
"""
import ctypes
import warnings

import numpy as np
import torch

from . import _capi
from . import lowering
from . import misc
from .interpolation_cubic import NaturalCubicSpline
from .interpolation_linear import LinearInterpolation

_METHODS = {"euler": _capi.EULER, "rk4": _capi.RK4_38, "dopri5": _capi.DOPRI5}
_ALL_TORCHDIFFEQ_METHODS = ("dopri8", "dopri5", "bosh3", "fehlberg2", "adaptive_heun", "euler", "midpoint", "rk4",
                            "explicit_adams", "implicit_adams", "fixed_adams", "scipy_solver")
_PRECISIONS = {"fp32": _capi.PREC_FP32, "bf16": _capi.PREC_BF16, "bf16x3": _capi.PREC_BF16X3}

_THIRD = 1 / 3
_TWO_THIRDS = 2 / 3

# default arithmetic of the final-layer tiles; override per call with options={'precision': ...}
default_precision = "fp32"


class FixedSchedule:
    """Host arrays describing the fixed grid, stage times and output map (kept alive while kernels are enqueued).

    Decreasing ``t`` follows torchdiffeq's _check_inputs (modules/torchdiffeq/torchdiffeq/_impl/misc.py:262-283): time is negated
    (tau = -t), the grid / stages / outputs are laid out in tau, and the vector field is evaluated at -tau with its sign flipped.
    Every step formula is linear in dt * k, so the kernels run the same arithmetic with the actual (decreasing) stage times and
    dt = -(tau1 - tau0).  ``perturb=True`` nudges the first stage time of a step to the next float and the last one (rk4) to the
    previous float, in the state's dtype and in tau, as _PerturbFunc does (misc.py:168-191, fixed_grid.py:9-29)."""

    def __init__(self, t_host, method, step_size, grid_constructor, func, z0, perturb=False):
        reverse = len(t_host) > 1 and bool(t_host[0] > t_host[1])
        t = -t_host if reverse else t_host
        if step_size is None:
            if grid_constructor is None:
                grid = t
            elif reverse:
                grid = -torch.as_tensor(grid_constructor(func, z0, -t))
            else:
                grid = grid_constructor(func, z0, t)
        else:
            if grid_constructor is not None:
                raise ValueError("step_size and grid_constructor are mutually exclusive arguments.")
            # solvers.py:77-88, same torch ops in t's dtype so the grid is bit-identical
            niters = torch.ceil((t[-1] - t[0]) / step_size + 1).item()
            grid = torch.arange(0, niters, dtype=t.dtype) * step_size + t[0]
            grid[-1] = t[-1]
        grid = torch.as_tensor(grid).detach().cpu()
        assert grid[0] == t[0] and grid[-1] == t[-1]
        g0, g1 = grid[:-1], grid[1:]
        dt = g1 - g0
        if method == "rk4":
            # rk_common.py:111-113; the cast to the state dtype is _PerturbFunc's (misc.py:181)
            stages = torch.stack([g0, g0 + dt * _THIRD, g0 + dt * _TWO_THIRDS, g1], dim=1)
        else:
            stages = g0.unsqueeze(1)
        self.n_steps = int(dt.numel())
        self.n_stages = int(stages.shape[1]) if self.n_steps else (4 if method == "rk4" else 1)
        stage_t = np.ascontiguousarray(stages.to(torch.float32).numpy())
        if perturb and self.n_steps:
            stage_t[:, 0] = np.nextafter(stage_t[:, 0], np.float32(np.inf))
            if method == "rk4":
                stage_t[:, -1] = np.nextafter(stage_t[:, -1], np.float32(-np.inf))
        dt32 = np.ascontiguousarray(dt.to(torch.float32).numpy())
        self.stage_t = np.ascontiguousarray(-stage_t) if reverse else stage_t
        self.dt = np.ascontiguousarray(-dt32) if reverse else dt32
        self.reverse = reverse
        # output map (solvers.py:106-117, 166-172), in tau
        T = int(t.numel())
        tn, gn = t.numpy(), grid.numpy()
        out_step = np.zeros(T, dtype=np.int64)
        out_mode = np.zeros(T, dtype=np.int32)
        out_slope = np.zeros(T, dtype=np.float32)
        if T > 1:
            step = np.searchsorted(gn[1:], tn[1:], side="left")
            a, b = gn[step], gn[step + 1]
            out_step[1:] = step
            mode = np.where(tn[1:] == a, 0, np.where(tn[1:] == b, 1, 2))
            out_mode[1:] = mode
            with np.errstate(all="ignore"):
                out_slope[1:] = ((tn[1:] - a) / (b - a)).astype(np.float32)
        self.out_step, self.out_mode, self.out_slope = out_step, out_mode, out_slope
        self.n_out = T


_SCHEDULE_CACHE = {}


def _schedule(t_host, method, options, func, z0):
    step_size = options.get("step_size")
    gc = options.get("grid_constructor")
    perturb = bool(options.get("perturb", False))
    if gc is None:
        key = (t_host.dtype, t_host.numpy().tobytes(), method, None if step_size is None else float(step_size), perturb)
        hit = _SCHEDULE_CACHE.get(key)
        if hit is None:
            if len(_SCHEDULE_CACHE) > 64:
                _SCHEDULE_CACHE.clear()
            hit = _SCHEDULE_CACHE[key] = FixedSchedule(t_host, method, step_size, None, func, z0, perturb)
        return hit
    return FixedSchedule(t_host, method, step_size, gc, func, z0, perturb)


def _np_ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


_VF_TYPES = {"matmul": _capi.VF_MATMUL, "evaluate": _capi.VF_EVALUATE, "derivative": _capi.VF_DERIVATIVE}


def _build_problem(X, spec, B, H, C, method, precision, sched):
    p = _capi.Problem()
    p.B, p.H, p.C = B, H, C
    p.vf_type = _VF_TYPES[getattr(spec, "vector_field_type", "matmul")]
    p.method = _METHODS[method]
    p.precision = precision
    m = p.mlp
    m.n_layers = len(spec.weights)
    keep = []
    for i, (w, b, a, s) in enumerate(zip(spec.weights, spec.biases, spec.acts, spec.slots)):
        wd = w.detach()
        if not wd.is_contiguous():
            wd = wd.contiguous()
        keep.append(wd)
        m.in_dim[i], m.out_dim[i], m.act[i], m.slot[i] = wd.shape[1], wd.shape[0], a, s
        m.W[i] = wd.data_ptr()
        if b is not None:
            bd = b.detach().contiguous()
            keep.append(bd)
            m.bias[i] = bd.data_ptr()
        else:
            m.bias[i] = None
    if getattr(spec, "gate", None) is not None:
        gw = spec.gate[0].detach().contiguous()
        keep.append(gw)
        m.W_gate = gw.data_ptr()
        if spec.gate[1] is not None:
            gbias = spec.gate[1].detach().contiguous()
            keep.append(gbias)
            m.bias_gate = gbias.data_ptr()
    if isinstance(X, LinearInterpolation):
        p.path.kind = _capi.PATH_LINEAR
        coeffs = X._coeffs.detach().contiguous()
        derivs = X._derivs.detach().contiguous()
        p.path.K = coeffs.size(-2)
        p.path.derivs = derivs.data_ptr()
        keep.append(derivs)
        if getattr(X, "gradient_matching_eps", None) is not None:   # SmoothLinearInterpolation (ncde_b200.interpolation)
            if coeffs.dtype != torch.float32:
                raise NotImplementedError("gradient matching inside cdeint needs float32 coefficients")
            if method == "dopri5":
                raise NotImplementedError("SmoothLinearInterpolation with method='dopri5' is not implemented")
            match = X.gradient_matching_coeffs.detach().contiguous()
            keep.append(match)
            p.path.match = match.data_ptr()
            p.path.match_terms = X.match_terms
            p.path.match_eps = float(X.gradient_matching_eps)
    else:
        p.path.kind = _capi.PATH_CUBIC
        coeffs = X._coeffs.detach().contiguous()
        p.path.K = coeffs.size(-2) + 1
        p.path.derivs = None
    knots = X._t.detach().to(torch.float32).contiguous()
    keep += [coeffs, knots]
    p.path.knots = knots.data_ptr()
    p.path.coeffs = coeffs.data_ptr()
    g = p.grid
    g.n_steps = sched.n_steps
    g.stage_t, g.dt = _np_ptr(sched.stage_t), _np_ptr(sched.dt)
    g.n_out = sched.n_out
    g.out_step, g.out_mode, g.out_slope = _np_ptr(sched.out_step), _np_ptr(sched.out_mode), _np_ptr(sched.out_slope)
    return p, keep


# counters for bench / tests: kernels enqueued by the last forward / backward call
last_launches = {"fwd": 0, "bwd": 0}


class _FixedSolve(torch.autograd.Function):
    """Forward: ncde_solve_fwd.  Backward: ncde_solve_bwd — the exact gradient of the discrete step loop, i.e. what
    autograd computes through the reference when adjoint=False."""

    @staticmethod
    def forward(ctx, X, spec, method, precision, sched, z0, coeffs_for_graph, *params):
        B, H = z0.shape
        C = spec.channels
        dev = z0.device
        problem, keep = _build_problem(X, spec, B, H, C, method, precision, sched)
        L = _capi.lib()
        need_grad = any(ctx.needs_input_grad[5:])
        z0c = z0.detach().contiguous()
        z_out = torch.empty(sched.n_out, B, H, dtype=torch.float32, device=dev)
        saved = None
        if need_grad:
            nbytes = L.ncde_solve_saved_bytes(ctypes.byref(problem), 1)
            saved = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        wbytes = L.ncde_solve_workspace_bytes(ctypes.byref(problem), 0)
        if wbytes == 0:
            # let the library produce the precise error
            _capi.check(L.ncde_solve_fwd(ctypes.byref(problem), z0c.data_ptr(), z_out.data_ptr(), None, 0, None, 0,
                                         None, None, None, _capi.stream_ptr(dev)))
        work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
        launches = ctypes.c_int64(0)
        _capi.check(L.ncde_solve_fwd(ctypes.byref(problem), z0c.data_ptr(), z_out.data_ptr(), _capi.ptr(saved),
                                     int(need_grad), work.data_ptr(), wbytes, None, None, ctypes.byref(launches),
                                     _capi.stream_ptr(dev)))
        last_launches["fwd"] = launches.value
        ctx.X, ctx.spec, ctx.method, ctx.precision, ctx.sched = X, spec, method, precision, sched
        ctx.saved_buf = saved
        ctx.shape = (B, H, C)
        ctx.coeffs_shape = coeffs_for_graph.shape
        ctx.n_params = len(params)
        return z_out

    @staticmethod
    def backward(ctx, grad_out):
        B, H, C = ctx.shape
        spec = ctx.spec
        dev = grad_out.device
        want_path_grad = ctx.needs_input_grad[6]
        problem, keep = _build_problem(ctx.X, spec, B, H, C, ctx.method, ctx.precision, ctx.sched)
        L = _capi.lib()
        g = grad_out.contiguous()
        grad_z0 = torch.empty(B, H, dtype=torch.float32, device=dev)
        # one gradient tensor per unique parameter, in the order the params were passed to apply()
        uniq = spec.unique_params
        grads = [torch.zeros_like(p, dtype=torch.float32, memory_format=torch.contiguous_format) for p, _, _ in uniq]
        by_layer_w, by_layer_b = {}, {}
        for gt, (_, kind, layer) in zip(grads, uniq):
            (by_layer_w if kind == "W" else by_layer_b)[layer] = gt
        n = len(spec.weights)
        gW = (ctypes.c_void_p * _capi.MAX_LAYERS)()
        gb = (ctypes.c_void_p * _capi.MAX_LAYERS)()
        first_of_slot = {}
        for i in range(n):
            j = first_of_slot.setdefault(spec.slots[i], i)
            gW[i] = by_layer_w[j].data_ptr()
            gb[i] = by_layer_b[j].data_ptr() if j in by_layer_b else None
        if getattr(spec, "gate", None) is not None:   # gate gradients: slot n_layers
            gW[n] = by_layer_w[n].data_ptr()
            gb[n] = by_layer_b[n].data_ptr() if n in by_layer_b else None
        # gradient w.r.t. the path coefficients (stacked CDEs, test_tricks.py:54-106): accumulated by the library into zeros
        grad_coeffs = torch.zeros_like(ctx.X._coeffs, memory_format=torch.contiguous_format) if want_path_grad else None
        wbytes = L.ncde_solve_workspace_bytes(ctypes.byref(problem), 2 if want_path_grad else 1)
        work = torch.empty(max(wbytes, 16), dtype=torch.uint8, device=dev)
        launches = ctypes.c_int64(0)
        _capi.check(L.ncde_solve_bwd(ctypes.byref(problem), g.data_ptr(), ctx.saved_buf.data_ptr(), grad_z0.data_ptr(),
                                     gW, gb, _capi.ptr(grad_coeffs), work.data_ptr(), wbytes, ctypes.byref(launches),
                                     _capi.stream_ptr(dev)))
        last_launches["bwd"] = launches.value
        # the stage records stay on ctx until autograd frees the graph: a second backward (retain_graph=True) is legal
        if want_path_grad:
            grad_coeffs = grad_coeffs.reshape(ctx.coeffs_shape)
        return (None, None, None, None, None, grad_z0, grad_coeffs) + tuple(grads)


class _AdjointFixedSolve(torch.autograd.Function):
    """adjoint=True with a fixed-grid method: forward without saving anything but the outputs, backward by integrating
    the augmented system (y, a, g_theta) backwards over every output interval (ncde_solve_adjoint_bwd) — the continuous
    adjoint of modules/torchdiffeq/torchdiffeq/_impl/adjoint.py:36-145."""

    @staticmethod
    def forward(ctx, X, spec, method, precision, sched, adj, t_host, z0, coeffs_for_graph, *params):
        B, H = z0.shape
        C = spec.channels
        dev = z0.device
        problem, keep = _build_problem(X, spec, B, H, C, method, precision, sched)
        L = _capi.lib()
        z0c = z0.detach().contiguous()
        z_out = torch.empty(sched.n_out, B, H, dtype=torch.float32, device=dev)
        wbytes = L.ncde_solve_workspace_bytes(ctypes.byref(problem), 0)
        work = torch.empty(max(wbytes, 16), dtype=torch.uint8, device=dev)
        launches = ctypes.c_int64(0)
        _capi.check(L.ncde_solve_fwd(ctypes.byref(problem), z0c.data_ptr(), z_out.data_ptr(), None, 0, work.data_ptr(),
                                     wbytes, None, None, ctypes.byref(launches), _capi.stream_ptr(dev)))
        last_launches["fwd"] = launches.value
        ctx.X, ctx.spec, ctx.precision, ctx.adj, ctx.t_host = X, spec, precision, adj, t_host
        ctx.shape = (B, H, C)
        ctx.save_for_backward(z_out)
        return z_out

    @staticmethod
    def backward(ctx, grad_out):
        (z_out,) = ctx.saved_tensors
        B, H, C = ctx.shape
        spec, t_host = ctx.spec, ctx.t_host
        adj_method, adj_options = ctx.adj
        dev = grad_out.device
        # backward schedule, last interval first, built in reversed time exactly like _check_inputs does (misc.py:262-283)
        T = int(t_host.numel())
        scheds = []
        for i in range(T - 1, 0, -1):
            tau = -t_host[i - 1:i + 1].flip(0)
            scheds.append(FixedSchedule(tau, adj_method, adj_options.get("step_size"), None, None, None))
        NS = 4 if adj_method == "rk4" else 1

        class _Grid:
            n_steps = sum(sc.n_steps for sc in scheds)
            n_out = T
            stage_t = np.ascontiguousarray(np.concatenate([-sc.stage_t.reshape(-1, NS) for sc in scheds]).astype(np.float32)) \
                if scheds else np.zeros((1, NS), dtype=np.float32)
            dt = np.ascontiguousarray(np.concatenate([sc.dt for sc in scheds]).astype(np.float32)) if scheds \
                else np.zeros(1, dtype=np.float32)
            out_step = np.zeros(1, dtype=np.int64)
            out_mode = np.zeros(1, dtype=np.int32)
            out_slope = np.zeros(1, dtype=np.float32)
        interval_steps = np.ascontiguousarray(np.array([sc.n_steps for sc in scheds] + [0], dtype=np.int64))
        problem, keep = _build_problem(ctx.X, spec, B, H, C, adj_method, ctx.precision, _Grid)
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
        if getattr(spec, "gate", None) is not None:   # gate gradients: slot n_layers
            n = len(spec.weights)
            gW[n] = by_layer_w[n].data_ptr()
            gb[n] = by_layer_b[n].data_ptr() if n in by_layer_b else None
        wbytes = L.ncde_solve_adjoint_workspace_bytes(ctypes.byref(problem))
        work = torch.empty(max(wbytes, 16), dtype=torch.uint8, device=dev)
        launches = ctypes.c_int64(0)
        _capi.check(L.ncde_solve_adjoint_bwd(ctypes.byref(problem), _np_ptr(interval_steps), T, z_out.data_ptr(),
                                             g.data_ptr(), grad_z0.data_ptr(), gW, gb, work.data_ptr(), wbytes,
                                             ctypes.byref(launches), _capi.stream_ptr(dev)))
        last_launches["bwd"] = launches.value
        ctx.nfe_bwd = _Grid.n_steps * NS
        return (None, None, None, None, None, None, None, grad_z0, None) + tuple(grads)


def cdeint(X, func, z0, t, adjoint=True, vector_field_type='matmul', **kwargs):
    r"""Solves z_t = z_{t_0} + \int_{t_0}^t f(s, z_s) dX_s on the GPU.

    Same signature, argument meaning, defaults and return layout ``(..., len(t), hidden_channels)`` as the
    reference's ``torchcde.cdeint`` (modules/torchcde/torchcde/solver.py:140-238).  ``kwargs`` are the torchdiffeq
    solver arguments (``method``, ``rtol``, ``atol``, ``options``, ``adjoint_*``).  One extra option is understood:
    ``options['precision']`` in {'fp32', 'bf16'} selects the arithmetic of the final-layer tiles.

    Differences, all loud: ``X`` must be a ``LinearInterpolation`` or ``NaturalCubicSpline`` of this package,
    ``func`` must lower to a Linear/activation chain (see ``lowering``), tensors must live on a CUDA device.
    ``vector_field_type`` 'matmul', 'evaluate' and 'derivative' are implemented (the latter two on the fixed-grid
    solvers).
    """
    if vector_field_type not in ['matmul', 'evaluate', 'derivative']:
        raise ValueError("vector_field_type string not recognised")
    if not isinstance(X, (LinearInterpolation, NaturalCubicSpline)):
        raise NotImplementedError("X must be a torchcde_b200 LinearInterpolation or NaturalCubicSpline")
    if not isinstance(z0, torch.Tensor):
        raise NotImplementedError("tuple state is not supported (it is dead code in the reference fork as well)")
    _capi.require_cuda(z0, X._coeffs)
    if z0.dtype != torch.float32 or X._coeffs.dtype != torch.float32:
        raise NotImplementedError("the fused solve runs an fp32 state; got z0 {} / coeffs {}".format(
            z0.dtype, X._coeffs.dtype))

    # solver.py:193-196
    if 'atol' not in kwargs:
        kwargs['atol'] = 1e-6
    if 'rtol' not in kwargs:
        kwargs['rtol'] = 1e-4
    method = kwargs.pop('method', None)
    options = dict(kwargs.pop('options', None) or {})
    if method is None:
        method = 'dopri5'  # misc.py:222-223
    if method not in _ALL_TORCHDIFFEQ_METHODS:
        raise ValueError('Invalid method "{}". Must be one of {}'.format(
            method, '{"' + '", "'.join(_ALL_TORCHDIFFEQ_METHODS) + '"}.'))
    if method not in _METHODS:
        raise NotImplementedError("method '{}' is not implemented by the fused solve (euler, rk4, dopri5 are)".format(method))
    precision = options.pop('precision', default_precision)
    if precision not in _PRECISIONS:
        raise ValueError("options['precision'] must be one of {}".format(sorted(_PRECISIONS)))

    if not isinstance(t, torch.Tensor):
        raise AssertionError("t must be a torch.Tensor")
    if not t.is_floating_point():
        raise TypeError('`t` must be a floating point Tensor but is a {}'.format(t.type()))
    assert t.ndimension() == 1, "t must be one dimensional"
    if t.requires_grad:
        raise NotImplementedError("gradients with respect to t are not implemented")
    t_host = misc.host_values(t)
    diff = t_host[1:] > t_host[:-1]
    assert bool(diff.all()) or bool((~diff).all()), 't must be strictly increasing or decreasing'
    decreasing = len(t_host) > 1 and bool(t_host[0] > t_host[1])
    if decreasing and (method == 'dopri5' or adjoint):
        raise NotImplementedError("decreasing integration times are implemented for the fixed-grid solvers (euler, rk4) with "
                                  "adjoint=False")

    batch_shape = z0.shape[:-1]
    H = z0.shape[-1]
    coeffs = X._coeffs
    if coeffs.shape[:-2] != batch_shape:
        raise ValueError("batch dimensions of z0 {} and of the control path {} differ".format(
            tuple(batch_shape), tuple(coeffs.shape[:-2])))
    C = coeffs.shape[-1] // (4 if isinstance(X, NaturalCubicSpline) else 1)
    spec = lowering.lower(func, H, C, vector_field_type)
    if vector_field_type != 'matmul':
        # f([z, X(t)]) / f([z, dX/dt(t)]) (solver.py:123-126): fixed-grid fp32 solves; gradients by backpropagation through the
        # steps or by the fixed-grid continuous adjoint
        if method == 'dopri5':
            raise NotImplementedError("vector_field_type='{}' is implemented for the fixed-grid solvers (euler, rk4) "
                                      "only".format(vector_field_type))
        if precision != 'fp32':
            raise NotImplementedError("vector_field_type='{}' runs in fp32 only".format(vector_field_type))
        if getattr(X, "gradient_matching_eps", None) is not None:
            raise NotImplementedError("vector_field_type='{}' is not implemented for gradient-matched paths".format(
                vector_field_type))
        if coeffs.requires_grad and torch.is_grad_enabled() and not adjoint:   # (with adjoint=True the path gets no gradient: below)
            raise NotImplementedError("gradients with respect to the control path need vector_field_type='matmul'")
    for w in spec.weights:
        _capi.require_cuda(w)
        if w.dtype != torch.float32:
            raise NotImplementedError("vector field parameters must be float32")
    if getattr(spec, "gate", None) is not None:
        # gated fields: fixed-grid fp32 solves (backpropagation through the steps or the fixed-grid continuous adjoint)
        if method == 'dopri5' or precision != 'fp32':
            raise NotImplementedError("gated vector fields are implemented for method euler / rk4 and precision fp32")
        if coeffs.requires_grad and torch.is_grad_enabled() and not adjoint:
            raise NotImplementedError("gradients with respect to the control path are not implemented for gated vector fields")

    if coeffs.requires_grad and torch.is_grad_enabled():
        if adjoint:
            # solver.py:201-221: path buffers that need gradients must be listed in adjoint_params, else they get none
            listed = [id(q) for q in (kwargs.get('adjoint_params') or ())]
            if any(id(buf) in listed for buf in X.buffers()):
                raise NotImplementedError("gradients with respect to the control path through the continuous adjoint "
                                          "are not implemented; use adjoint=False")
            warnings.warn("One of the inputs to the control path X requires gradients but is not listed in "
                          "`options['adjoint_params']`. This is probably a mistake: it will not receive a gradient "
                          "when using the adjoint vector_field_type. Either have the input not require gradients (if "
                          "that was unintended), or include it (and every other parameter needing gradients) in "
                          "`adjoint_params`.")
            coeffs = coeffs.detach()
        else:
            if precision != 'fp32':
                raise NotImplementedError("gradients with respect to the control path need options['precision']='fp32'")
            if getattr(X, "gradient_matching_eps", None) is not None:
                raise NotImplementedError("gradients with respect to a gradient-matched (smoothed) control path are "
                                          "not implemented")
    if method == 'dopri5':
        from . import adaptive
        out = adaptive.solve(X, func, spec, z0, t, t_host, adjoint, options, kwargs, coeffs=coeffs)
    else:
        unused = {k: v for k, v in options.items() if k not in ('step_size', 'grid_constructor', 'perturb', 'interp')}
        if options.get('interp', 'linear') != 'linear':
            raise NotImplementedError("only interp='linear' is implemented for fixed-grid output")
        if options.get('perturb', False) and adjoint:
            raise NotImplementedError("perturb=True is implemented with adjoint=False")
        if unused:
            warnings.warn('{}: Unexpected arguments {}'.format('RK4' if method == 'rk4' else 'Euler', unused))
        z0f = z0.reshape(-1, H)
        sched = _schedule(t_host, method, options, func, z0f)
        params = [p for p, _, _ in spec.unique_params]
        Xf = X
        if len(batch_shape) != 1:
            Xf = _flatten_path(X)
        if adjoint:
            # adjoint.py:159-171: adjoint_* default to the forward method / options (minus `norm`)
            adj_method = kwargs.get('adjoint_method') or method
            adj_options = kwargs.get('adjoint_options')
            if adj_options is None:
                adj_options = {k: v for k, v in options.items() if k != 'norm'}
            if adj_method not in ('euler', 'rk4'):
                raise NotImplementedError("adjoint_method '{}' is not implemented (euler, rk4 are)".format(adj_method))
            if adj_options.get('grid_constructor') is not None:
                raise NotImplementedError("grid_constructor is not implemented for the adjoint pass")
            out = _AdjointFixedSolve.apply(Xf, spec, method, _PRECISIONS[precision], sched, (adj_method, dict(adj_options)),
                                           t_host, z0f, coeffs, *params)
        else:
            out = _FixedSolve.apply(Xf, spec, method, _PRECISIONS[precision], sched, z0f, coeffs, *params)
        if hasattr(func, "nfe"):
            func.nfe += sched.n_steps * sched.n_stages
        out = out.reshape(sched.n_out, *batch_shape, H)

    # solver.py:227-229: (T, ..., H) -> (..., T, H)
    batch_dims = range(1, len(out.shape) - 1)
    return out.permute(*batch_dims, 0, -1)


def _flatten_path(X):
    """View of X with all leading batch dimensions merged into one."""
    coeffs = X._coeffs.reshape(-1, *X._coeffs.shape[-2:])
    return type(X)(coeffs, misc.attach_host(X._t, X._t_host))
