"""Prints the tiling plan the library picks for a shape (no GPU needed): python tools/plan_probe.py B H C HH n_layers [fp32|bf16]"""
import ctypes, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "online-neural-cdes_b200")]
os.environ["NCDE_DEBUG_PLAN"] = "1"
from torchcde_b200 import _capi


def probe(B, H, C, HH, n, prec="bf16"):
    L = _capi.lib()
    P = _capi.Problem()
    P.B, P.H, P.C, P.method = B, H, C, _capi.RK4_38
    P.precision = _capi.PREC_BF16 if prec == "bf16" else _capi.PREC_FP32
    dims = [(H, HH)] + [(HH, HH)] * (n - 1) + [(HH, H * C)]
    P.mlp.n_layers = len(dims)
    for i, (a, b) in enumerate(dims):
        P.mlp.in_dim[i], P.mlp.out_dim[i] = a, b
        P.mlp.act[i] = _capi.ACT_TANH if i == len(dims) - 1 else _capi.ACT_RELU
        P.mlp.slot[i] = i
        P.mlp.W[i] = 16  # never dereferenced by the planner
    L.ncde_solve_workspace_bytes.restype = ctypes.c_size_t
    return L.ncde_solve_workspace_bytes(ctypes.byref(P), 1)


if __name__ == "__main__":
    a = sys.argv[1:]
    probe(*[int(v) for v in a[:5]], *(a[5:6]))
