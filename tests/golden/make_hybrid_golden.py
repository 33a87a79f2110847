"""Mints tests/golden/hybrid.pt from the REAL reference (src/ncde/interpolation.py loaded by path; run in the build container:
    PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference/modules/torchcde:/root/reference/modules/torchdiffeq python tests/golden/make_hybrid_golden.py
"""
import importlib.util
import os
import torch

spec = importlib.util.spec_from_file_location("ref_interpolation", "/root/reference/src/ncde/interpolation.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

torch.manual_seed(0)
cases = []
for B, L, C, rect, p_obs in [(4, 8, 5, [2, 3, 4], 0.3), (6, 20, 9, [4, 5, 6, 7, 8], 0.15), (3, 12, 4, [3], 0.5), (5, 10, 3, [1, 2], 0.2)]:
    x = torch.randn(B, L, C)
    x[..., 0] = torch.arange(L, dtype=torch.float32)
    miss = torch.rand(B, L, C) > p_obs
    miss[..., 0] = False
    for c in range(1, C):
        if c not in rect:
            miss[..., c] = torch.rand(B, L) > 0.8
    x[miss] = float("nan")
    out = ref._prepare_linear_rectilinear_hybrid(x.clone(), rectilinear_indices=rect)
    cases.append({"x": x, "rect": rect, "out": out})
torch.save(cases, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hybrid.pt"))
print([tuple(c["out"].shape) for c in cases])
