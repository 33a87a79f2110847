"""CPU-side checks: the C-ABI library loads and exports every declared symbol; the host schedule matches the
reference's grid construction; vector-field lowering; shard arithmetic.  No GPU needed."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import cde_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    from torchcde_b200 import _capi
    if not os.path.exists(_capi.LIB_PATH):
        g.build()
    return _capi.lib()


def test_library_exports_every_declared_symbol(built):
    from torchcde_b200 import _capi
    header = open(os.path.join(ROOT, "include", "ncde_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(ncde_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(built, name), "library does not export " + name
    assert sorted(_capi.SYMBOLS) == declared
    assert built.ncde_abi_version() == 4
    assert b"sm_100a" in built.ncde_version()


def test_struct_layout_matches_header(built):
    """sizeof(ncde_problem_t) as compiled by gcc == ctypes mirror (guards against field drift)."""
    import subprocess
    import tempfile
    from torchcde_b200 import _capi
    src = '#include <stdio.h>\n#include "ncde_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu", sizeof(ncde_problem_t),' \
          'sizeof(ncde_mlp_t), sizeof(ncde_path_t), sizeof(ncde_fixed_grid_t), sizeof(ncde_adaptive_t));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(v) for v in subprocess.check_output([exe]).split()]
    import ctypes
    assert sizes == [ctypes.sizeof(_capi.Problem), ctypes.sizeof(_capi.Mlp), ctypes.sizeof(_capi.Path),
                     ctypes.sizeof(_capi.FixedGrid), ctypes.sizeof(_capi.Adaptive)]


def test_invalid_arguments_are_rejected_without_a_gpu(built):
    from torchcde_b200 import _capi
    assert built.ncde_forward_fill(0, None, None, 1, 2, 3, None) == _capi.ERR_INVALID
    assert b"forward_fill" in built.ncde_last_error()
    assert built.ncde_rectilinear_prepare(0, 1, 1, 1, 2, 3, 7, None, None) == _capi.ERR_INVALID
    assert built.ncde_natural_cubic_coeffs(0, 1, 1, 1, 1, 1, 3, 1, 1, None) == _capi.ERR_INVALID
    # entry points added with the widened rows: argument validation happens before any CUDA call
    assert built.ncde_ragged_interpolate(0, 0, None, None, None, 1, 4, 3, 0, 1, 0, 0, None, None, None) == _capi.ERR_INVALID
    assert built.ncde_ragged_interpolate(7, 0, 1, 1, 1, 1, 4, 3, 0, 1, 0, 0, 1, None, None) == _capi.ERR_INVALID
    assert b"unknown method" in built.ncde_last_error()
    assert built.ncde_ragged_interpolate(1, 0, 1, 1, 1, 1, 4, 3, 5, 1, 0, 0, 1, None, None) == _capi.ERR_INVALID   # time index 5 of 3 channels
    assert built.ncde_ragged_interpolate(0, 0, 1, 1, 1, 1, 4, 3, 0, 1, 1, 0, 1, None, None) == _capi.ERR_INVALID   # intensity needs rectilinear
    assert built.ncde_ragged_scratch_bytes(2, 0, 10, 8, 3) > built.ncde_ragged_scratch_bytes(0, 0, 10, 8, 3) > 0
    assert built.ncde_path_eval_bwd(0, 0, None, 1, 4, 3, None, 1, 0, None, None, None) == _capi.ERR_INVALID
    assert built.ncde_path_eval_bwd(5, 0, 1, 1, 4, 3, 1, 1, 0, 1, 1, None) == _capi.ERR_INVALID
    with pytest.raises(ValueError):
        _capi.check(_capi.ERR_INVALID)


def test_problem_validation_without_a_gpu(built):
    """make_plan rejects malformed / unsupported problems before touching the device: the workspace query returns 0 and the
    solve entry point reports why (mapped to ValueError / NotImplementedError by _capi.check)."""
    import ctypes
    from torchcde_b200 import _capi

    def problem(H=8, C=3, HH=8, vf=0, gated=False, precision=0, method=1, act0=1):
        p = _capi.Problem()
        p.B, p.H, p.C, p.method, p.precision, p.vf_type = 4, H, C, method, precision, vf
        d0 = H + (C if vf else 0)
        out = H if vf else H * C
        m = p.mlp
        m.n_layers = 2
        m.in_dim[0], m.out_dim[0], m.act[0] = d0, HH, act0
        m.in_dim[1], m.out_dim[1], m.act[1] = HH, out, _capi.ACT_TANH
        m.W[0] = m.W[1] = 1          # never dereferenced by the checks exercised here
        if gated:
            m.W_gate = 1
        p.grid.n_steps = 0
        return p

    ok = problem()
    assert built.ncde_solve_workspace_bytes(ctypes.byref(ok), 0) > 0
    assert built.ncde_solve_workspace_bytes(ctypes.byref(ok), 1) > 0
    assert built.ncde_solve_workspace_bytes(ctypes.byref(ok), 2) > 0          # with the path gradient
    assert built.ncde_solve_workspace_bytes(ctypes.byref(problem(vf=1)), 1) > 0
    assert built.ncde_solve_workspace_bytes(ctypes.byref(problem(vf=1)), 2) == 0   # path gradient needs matmul
    assert built.ncde_solve_workspace_bytes(ctypes.byref(problem(gated=True)), 1) > 0
    assert built.ncde_solve_adjoint_workspace_bytes(ctypes.byref(problem(vf=2, gated=True))) > 0
    for bad, exc in [(problem(vf=1, precision=1), NotImplementedError),      # evaluate mode on bf16 tiles
                     (problem(gated=True, precision=1), NotImplementedError),
                     (problem(vf=3), ValueError),
                     (problem(HH=300), NotImplementedError),                   # final-layer input wider than 256
                     (problem(act0=_capi.ACT_GATE_IN), ValueError)]:          # gate-in layer needs out_dim == 2 * in_dim
        assert built.ncde_solve_workspace_bytes(ctypes.byref(bad), 0) == 0
        rc = built.ncde_solve_fwd(ctypes.byref(bad), 1, 1, None, 0, None, 0, None, None, None, None)
        with pytest.raises(exc):
            _capi.check(rc)
    # adaptive / adaptive-adjoint entry points refuse the fixed-grid-only modes
    assert built.ncde_solve_adjoint_adaptive_workspace_bytes(ctypes.byref(problem(vf=1, method=2))) == 0


def test_product_refuses_cpu_tensors(built):
    import torchcde_b200 as tc
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tc.linear_interpolation_coeffs(torch.zeros(2, 3, 2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tc.natural_cubic_coeffs(torch.zeros(2, 3, 2))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("step_size", [1, 0.5, 0.3, None])
@pytest.mark.parametrize("method", ["rk4", "euler"])
def test_schedule_matches_reference_grid(dtype, step_size, method):
    from torchcde_b200.solver import FixedSchedule
    g = torch.Generator().manual_seed(3)
    t = torch.cat([torch.zeros(1), torch.rand(6, generator=g).cumsum(0) * 2.1]).to(dtype)
    sched = FixedSchedule(t, method, step_size, None, None, None)
    grid = O.fixed_grid(t, step_size)
    assert sched.n_steps == len(grid) - 1
    # stage times and dt as the oracle's step loop computes them
    for s, (a, b) in enumerate(zip(grid[:-1], grid[1:])):
        dt = b - a
        assert sched.dt[s] == np.float32(dt.to(torch.float32).item())
        if method == "rk4":
            want = [a, a + dt * (1 / 3), a + dt * (2 / 3), b]
        else:
            want = [a]
        got = sched.stage_t[s]
        assert [np.float32(w.to(torch.float32).item()) for w in want] == list(got)
    # output map: replay the reference loop (solvers.py:106-117)
    j = 1
    for s, (a, b) in enumerate(zip(grid[:-1], grid[1:])):
        while j < len(t) and b >= t[j]:
            assert sched.out_step[j] == s
            if t[j] == a:
                assert sched.out_mode[j] == 0
            elif t[j] == b:
                assert sched.out_mode[j] == 1
            else:
                assert sched.out_mode[j] == 2
                assert sched.out_slope[j] == np.float32(((t[j] - a) / (b - a)).to(torch.float32).item())
            j += 1
    assert j == len(t)


def test_lowering_recognises_reference_fields():
    from torchcde_b200 import lowering, _capi
    f = O.SharedMLPField(5, 8, 12, 3)
    spec = lowering.lower(f, 8, 5)
    assert [tuple(w.shape) for w in spec.weights] == [(12, 8), (12, 12), (12, 12), (40, 12)]
    assert spec.acts == [_capi.ACT_RELU, _capi.ACT_RELU, _capi.ACT_RELU, _capi.ACT_TANH]
    assert spec.slots == [0, 1, 1, 2]          # the middle Linear is one shared object
    assert len(spec.unique_params) == 6
    assert f.nfe == 0                           # the validation probe does not count as an evaluation
    toy = O.ToyField(2, 32, width=128)
    spec = lowering.lower(toy, 32, 2)           # via torch.fx
    assert [tuple(w.shape) for w in spec.weights] == [(32, 32), (128, 32), (64, 128)]
    assert spec.acts == [_capi.ACT_RELU, _capi.ACT_RELU, _capi.ACT_TANH]


def test_lowering_rejects_what_it_cannot_run():
    from torchcde_b200 import lowering

    class Sig(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.v = torch.nn.Parameter(torch.rand(1, 1, 3))

        def forward(self, t, z):
            return z.sigmoid().unsqueeze(-1) + self.v

    with pytest.raises(NotImplementedError):
        lowering.lower(Sig(), 4, 3)
    with pytest.raises(ValueError):
        lowering.lower(O.SharedMLPField(5, 8, 12, 2), 8, 4)


def test_shard_bounds_cover_batch():
    from torchcde_b200.distributed import shard_bounds
    for n in (1, 7, 8, 1024, 8191):
        for w in (1, 2, 3, 8):
            cuts = [shard_bounds(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
