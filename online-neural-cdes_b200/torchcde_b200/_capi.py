"""ctypes binding of libncde_b200.so (the C ABI declared in include/ncde_b200.h).

The product path has no CPU fallback: if the shared library is missing, or a tensor is not on a CUDA device, every
entry point raises.  PyTorch is used for device memory and streams only.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libncde_b200.so")

MAX_LAYERS = 8
MAX_STAGES = 7

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_WORKSPACE = 0, -1, -2, -3, -4
F32, F64 = 0, 1
PATH_LINEAR, PATH_CUBIC = 0, 1
EULER, RK4_38, DOPRI5 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_TANH, ACT_GATE_IN = 0, 1, 2, 3
PREC_FP32, PREC_BF16, PREC_BF16X3 = 0, 1, 2
VF_MATMUL, VF_EVALUATE, VF_DERIVATIVE = 0, 1, 2
RAGGED_LINEAR, RAGGED_RECTILINEAR, RAGGED_CUBIC = 0, 1, 2
FLAG_NAN_TIME, FLAG_NONFINITE, FLAG_DT_UNDERFLOW, FLAG_MAX_STEPS = 1, 2, 4, 8

c_float_p = ctypes.POINTER(ctypes.c_float)


class Mlp(ctypes.Structure):
    _fields_ = [("n_layers", ctypes.c_int32),
                ("in_dim", ctypes.c_int32 * MAX_LAYERS),
                ("out_dim", ctypes.c_int32 * MAX_LAYERS),
                ("act", ctypes.c_int32 * MAX_LAYERS),
                ("slot", ctypes.c_int32 * MAX_LAYERS),
                ("W", ctypes.c_void_p * MAX_LAYERS),
                ("bias", ctypes.c_void_p * MAX_LAYERS),
                ("W_gate", ctypes.c_void_p),
                ("bias_gate", ctypes.c_void_p)]


class Path(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32),
                ("K", ctypes.c_int64),
                ("knots", ctypes.c_void_p),
                ("coeffs", ctypes.c_void_p),
                ("derivs", ctypes.c_void_p),
                ("match", ctypes.c_void_p),
                ("match_terms", ctypes.c_int32),
                ("match_eps", ctypes.c_float)]


class FixedGrid(ctypes.Structure):
    _fields_ = [("n_steps", ctypes.c_int64),
                ("stage_t", ctypes.c_void_p),
                ("dt", ctypes.c_void_p),
                ("n_out", ctypes.c_int64),
                ("out_step", ctypes.c_void_p),
                ("out_mode", ctypes.c_void_p),
                ("out_slope", ctypes.c_void_p)]


class Adaptive(ctypes.Structure):
    _fields_ = [("rtol", ctypes.c_double), ("atol", ctypes.c_double), ("min_step", ctypes.c_double),
                ("max_step", ctypes.c_double), ("first_step", ctypes.c_double), ("safety", ctypes.c_double),
                ("ifactor", ctypes.c_double), ("dfactor", ctypes.c_double),
                ("max_attempts", ctypes.c_int64), ("n_out", ctypes.c_int64), ("out_t", ctypes.c_void_p)]


class Problem(ctypes.Structure):
    _fields_ = [("B", ctypes.c_int64), ("H", ctypes.c_int32), ("C", ctypes.c_int32),
                ("method", ctypes.c_int32), ("precision", ctypes.c_int32),
                ("mlp", Mlp), ("path", Path), ("grid", FixedGrid), ("adaptive", Adaptive),
                ("vf_type", ctypes.c_int32)]


_lib = None

# every symbol include/ncde_b200.h declares; tests check that the library exports all of them
SYMBOLS = ["ncde_version", "ncde_last_error", "ncde_abi_version", "ncde_forward_fill", "ncde_rectilinear_prepare",
           "ncde_linear_fill_missing", "ncde_cubic_scratch_bytes", "ncde_natural_cubic_coeffs", "ncde_linear_derivs",
           "ncde_path_eval", "ncde_ragged_scratch_bytes", "ncde_ragged_interpolate", "ncde_path_eval_bwd", "ncde_logsig_windows", "ncde_hybrid_compact", "ncde_smooth_matching_coeffs", "ncde_path_eval_smooth", "ncde_solve_saved_bytes", "ncde_solve_workspace_bytes", "ncde_solve_fwd",
           "ncde_solve_bwd", "ncde_solve_adaptive_fwd", "ncde_solve_adjoint_workspace_bytes",
           "ncde_solve_adjoint_bwd", "ncde_solve_adjoint_adaptive_workspace_bytes",
           "ncde_solve_adjoint_adaptive_bwd", "ncde_profile_enable", "ncde_profile_read"]

PROF_CLASSES = ["hidden_fwd", "field_fwd", "field_bwd", "hidden_bwd", "hidden_wgrad", "other", "solve_fwd", "solve_bwd"]


def lib():
    """Load the shared library once.  Raises if it has not been built (``python -c 'import __graft_entry__ as g;
    g.build()'`` or ``make -C online-neural-cdes_b200/csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libncde_b200.so is not built ({}); there is no CPU or PyTorch fallback. Build it with "
                           "`make -C online-neural-cdes_b200/csrc`.".format(LIB_PATH))
    L = ctypes.CDLL(LIB_PATH)
    vp, i64, i32, sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_size_t
    L.ncde_version.restype = ctypes.c_char_p
    L.ncde_last_error.restype = ctypes.c_char_p
    L.ncde_abi_version.restype = i32
    L.ncde_forward_fill.argtypes = [i32, vp, vp, i64, i64, i64, vp]
    L.ncde_rectilinear_prepare.argtypes = [i32, vp, vp, i64, i64, i64, i32, vp, vp]
    L.ncde_linear_fill_missing.argtypes = [i32, vp, vp, i64, i64, i64, vp]
    L.ncde_cubic_scratch_bytes.argtypes = [i32, i64, i64, i64]
    L.ncde_cubic_scratch_bytes.restype = sz
    L.ncde_natural_cubic_coeffs.argtypes = [i32, vp, vp, vp, i64, i64, i64, i32, vp, vp]
    L.ncde_linear_derivs.argtypes = [i32, vp, vp, vp, i64, i64, i64, vp]
    L.ncde_path_eval.argtypes = [i32, i32, vp, vp, vp, i64, i64, i64, vp, i64, i32, vp, vp, vp]
    L.ncde_ragged_scratch_bytes.argtypes = [i32, i32, i64, i64, i64]
    L.ncde_ragged_scratch_bytes.restype = sz
    L.ncde_ragged_interpolate.argtypes = [i32, i32, vp, vp, vp, i64, i64, i64, i32, i32, i32, i32, vp, vp, vp]
    L.ncde_path_eval_bwd.argtypes = [i32, i32, vp, i64, i64, i64, vp, i64, i32, vp, vp, vp]
    L.ncde_logsig_windows.argtypes = [i32, vp, vp, vp, vp, i64, i64, i32, i32, i32, vp]
    L.ncde_hybrid_compact.argtypes = [i32, vp, vp, vp, vp, i64, i64, i64, vp]
    L.ncde_smooth_matching_coeffs.argtypes = [i32, vp, vp, i64, i64, i64, ctypes.c_double, i32, vp]
    L.ncde_path_eval_smooth.argtypes = [i32, vp, vp, vp, vp, i32, ctypes.c_double, i64, i64, i64, vp, i64, i32, vp, vp]
    L.ncde_solve_saved_bytes.argtypes = [ctypes.POINTER(Problem), i32]
    L.ncde_solve_saved_bytes.restype = sz
    L.ncde_solve_workspace_bytes.argtypes = [ctypes.POINTER(Problem), i32]
    L.ncde_solve_workspace_bytes.restype = sz
    L.ncde_solve_fwd.argtypes = [ctypes.POINTER(Problem), vp, vp, vp, i32, vp, sz, vp, vp,
                                 ctypes.POINTER(ctypes.c_int64), vp]
    L.ncde_solve_bwd.argtypes = [ctypes.POINTER(Problem), vp, vp, vp, ctypes.POINTER(vp), ctypes.POINTER(vp), vp, vp,
                                 sz, ctypes.POINTER(ctypes.c_int64), vp]
    L.ncde_solve_adaptive_fwd.argtypes = [ctypes.POINTER(Problem), vp, vp, vp, sz, vp, ctypes.POINTER(ctypes.c_int64), vp]
    L.ncde_solve_adjoint_workspace_bytes.argtypes = [ctypes.POINTER(Problem)]
    L.ncde_solve_adjoint_workspace_bytes.restype = sz
    L.ncde_solve_adjoint_bwd.argtypes = [ctypes.POINTER(Problem), vp, i64, vp, vp, vp, ctypes.POINTER(vp), ctypes.POINTER(vp),
                                         vp, sz, ctypes.POINTER(ctypes.c_int64), vp]
    L.ncde_solve_adjoint_adaptive_workspace_bytes.argtypes = [ctypes.POINTER(Problem)]
    L.ncde_solve_adjoint_adaptive_workspace_bytes.restype = sz
    L.ncde_solve_adjoint_adaptive_bwd.argtypes = [ctypes.POINTER(Problem), vp, vp, vp, ctypes.POINTER(vp), ctypes.POINTER(vp),
                                                  vp, sz, vp, ctypes.POINTER(ctypes.c_int64), vp]
    L.ncde_profile_enable.argtypes = [i32]
    L.ncde_profile_read.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)]
    for name in SYMBOLS:
        fn = getattr(L, name)
        if fn.restype is ctypes.c_int and name not in ("ncde_abi_version",):
            fn.restype = i32
    if L.ncde_abi_version() != 4:
        raise RuntimeError("libncde_b200.so ABI version mismatch")
    _lib = L
    return L


class NcdeError(RuntimeError):
    pass


def check(rc):
    """Map C status codes to the exceptions the reference raises (SURVEY §8b error conventions)."""
    if rc == OK:
        return
    msg = lib().ncde_last_error().decode()
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise NcdeError("libncde_b200 error {}: {}".format(rc, msg))


def dtype_code(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.float64:
        return F64
    raise ValueError("only float32 / float64 tensors are supported, got {}".format(t.dtype))


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("torchcde_b200 runs on CUDA (B200) only and has no CPU fallback; got a tensor on "
                               "{}".format(t.device))


def stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def profile_enable(classes):
    """Bracket every launch of the named kernel classes with CUDA events (bench.py's live kernel timing)."""
    mask = 0
    for c in classes:
        mask |= 1 << PROF_CLASSES.index(c)
    check(lib().ncde_profile_enable(mask))


def profile_read():
    """-> {class: (total_ms, launches)}; synchronises on the recorded events and clears them."""
    ms = (ctypes.c_double * len(PROF_CLASSES))()
    cnt = (ctypes.c_int64 * len(PROF_CLASSES))()
    check(lib().ncde_profile_read(ms, cnt))
    return {c: (ms[i], cnt[i]) for i, c in enumerate(PROF_CLASSES)}
