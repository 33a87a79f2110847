"""Mint golden vectors by running the REAL reference (vendored torchcde 0.2.0 / torchdiffeq 0.2.1, imported
read-only from /root/reference) on seeded inputs.  Run in the authoring container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Outputs tests/golden/*.pt (small).  Nothing at test / bench time reads /root/reference; tests read these files.
"""
import importlib.util
import os
import sys
import types
import warnings

import torch

REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(REF, "modules/torchcde"))
sys.path.insert(0, os.path.join(REF, "modules/torchdiffeq"))
import torchcde  # noqa: E402
import torchdiffeq  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# reference vector field (src/ncde/vector_fields/base.py has no external deps)
_vf = _load(os.path.join(REF, "src/ncde/vector_fields/base.py"), "ref_vf_base")
# toy CDEFunc: stub matplotlib (absent) so the experiment file imports
_mpl = types.ModuleType("matplotlib")
_plt = types.ModuleType("matplotlib.pyplot")
_mpl.pyplot = _plt
sys.modules.setdefault("matplotlib", _mpl)
sys.modules.setdefault("matplotlib.pyplot", _plt)
_toy = _load(os.path.join(REF, "experiments/sim_bm_toy_example.py"), "ref_toy")


def save(name, obj):
    path = os.path.join(HERE, name + ".pt")
    torch.save(obj, path)
    print("wrote", path, os.path.getsize(path), "bytes")


def nan_mask_(x, p, gen, keep_first=True, skip_channel=None):
    drop = torch.rand(x.shape, generator=gen) < p
    if keep_first:
        drop[..., 0, :] = False
    if skip_channel is not None:
        drop[..., skip_channel] = False
    x[drop] = float("nan")
    return x


# ----------------------------------------------------------------------------------------------------------------
def golden_interpolation():
    g = torch.Generator().manual_seed(1234)
    out = {}
    nan = float("nan")
    # the reference's own hand-written rectilinear case (modules/torchcde/test/test_linear_interpolation.py:124-145)
    t1 = torch.tensor([0.1, 0.2, 0.9]).view(-1, 1)
    t2 = torch.tensor([0.2, 0.3]).view(-1, 1)
    x1 = torch.tensor([0.4, nan, 1.1]).view(-1, 1)
    x2 = torch.tensor([nan, 2.]).view(-1, 1)
    x = torch.nn.utils.rnn.pad_sequence([torch.cat((t1, x1), -1), torch.cat((t2, x2), -1)], batch_first=True,
                                        padding_value=nan)
    x[:, :, 0] = torchcde.misc.forward_fill(x[:, :, 0], fill_index=-1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out["rect_hand"] = {"x": x.clone(), "time_index": 0,
                            "out": torchcde.linear_interpolation_coeffs(x.clone(), rectilinear=0)}

    cases = []
    for (B, L, C, dtype, p, tidx) in [(5, 9, 4, torch.float32, 0.3, 0), (3, 17, 6, torch.float64, 0.5, 2),
                                      (4, 6, 3, torch.float32, 0.0, 0), (2, 30, 11, torch.float32, 0.7, 0)]:
        x = torch.randn(B, L, C, generator=g, dtype=dtype)
        x[..., tidx] = torch.rand(B, L, generator=g, dtype=dtype).cumsum(-1)
        nan_mask_(x, p, g, keep_first=False, skip_channel=tidx)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ffill = torchcde.misc.forward_fill(x.clone())
            rect_raw = torchcde.interpolation_linear._prepare_rectilinear_interpolation(x.clone(), tidx)
            rect = torchcde.linear_interpolation_coeffs(x.clone(), rectilinear=tidx)
            x0 = x.clone()
            rect_init = torchcde.linear_interpolation_coeffs(x0, rectilinear=tidx, initial_value_if_nan=0.25)
        cases.append({"x": x, "time_index": tidx, "ffill": ffill, "rect_raw": rect_raw, "rect": rect,
                      "rect_init": rect_init, "x_after_init": x0})
    out["rect_random"] = cases

    # linear coeffs with interior NaN fill, with and without explicit t
    lin = []
    for (shape, dtype, p, with_t) in [((4, 12, 3), torch.float32, 0.4, False), ((2, 3, 9, 2), torch.float64, 0.3, True),
                                      ((7, 5), torch.float32, 0.5, True), ((3, 8, 4), torch.float32, 0.0, False)]:
        x = torch.randn(*shape, generator=g, dtype=dtype)
        nan_mask_(x, p, g, keep_first=False)
        x[..., 0, 0] = float("nan") if p > 0 else x[..., 0, 0]  # leading NaN somewhere
        if len(shape) >= 3:
            x[0, ..., -1] = float("nan")  # an all-NaN channel in one series -> zeros
        t = torch.rand(shape[-2], generator=g, dtype=dtype).cumsum(0) if with_t else None
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            c = torchcde.linear_interpolation_coeffs(x.clone(), t)
            cf = torchcde.linear_interpolation_coeffs(x.clone(), t, forward_fill=True)
        lin.append({"x": x, "t": t, "coeffs": c, "coeffs_ffill": cf})
    out["linear_coeffs"] = lin

    # natural cubic coefficients
    cub = []
    for (shape, dtype, p, with_t) in [((4, 10, 3), torch.float32, 0.0, False), ((2, 2, 14, 2), torch.float64, 0.0, True),
                                      ((3, 9, 4), torch.float64, 0.35, True), ((5, 2, 3), torch.float32, 0.0, False),
                                      ((3, 161, 5), torch.float32, 0.0, False), ((2, 7, 3), torch.float32, 0.4, False)]:
        x = torch.randn(*shape, generator=g, dtype=dtype)
        nan_mask_(x, p, g, keep_first=False)
        t = torch.rand(shape[-2], generator=g, dtype=dtype).cumsum(0) if with_t else None
        cub.append({"x": x, "t": t, "coeffs_v1": torchcde.natural_cubic_coeffs(x.clone(), t),
                    "coeffs_v0": torchcde.natural_spline_coeffs(x.clone(), t) if hasattr(torchcde, "natural_spline_coeffs")
                    else torchcde.natural_cubic_spline_coeffs(x.clone(), t)})
    out["cubic_coeffs"] = cub

    # evaluate / derivative / knot indices at probe times (incl. exact knots, outside the interval)
    ev = []
    for (B, K, C, dtype, with_t) in [(3, 8, 4, torch.float32, False), (2, 13, 3, torch.float64, True),
                                     (4, 5, 2, torch.float32, True)]:
        x = torch.randn(B, K, C, generator=g, dtype=dtype)
        t = torch.rand(K, generator=g, dtype=dtype).cumsum(0) if with_t else None
        knots = t if with_t else torch.linspace(0, K - 1, K, dtype=dtype)
        lo, hi = knots[0], knots[-1]
        probes = torch.cat([knots, lo - 0.5 + torch.zeros(1, dtype=dtype), hi + 0.7 + torch.zeros(1, dtype=dtype),
                            lo + (hi - lo) * torch.rand(12, generator=g, dtype=dtype),
                            knots[1:-1] + (knots[1:-1] * 1e-7)])
        lin_c = torchcde.linear_interpolation_coeffs(x, t)
        cub_c = torchcde.natural_cubic_coeffs(x, t)
        LX = torchcde.LinearInterpolation(lin_c, t)
        CX = torchcde.NaturalCubicSpline(cub_c, t)
        rec = {"x": x, "t": t, "probes": probes, "lin_coeffs": lin_c, "cub_coeffs": cub_c,
               "lin_index": torch.stack([LX._interpret_t(p)[1] for p in probes]),
               "cub_index": torch.stack([CX._interpret_t(p)[1] for p in probes]),
               "lin_eval": torch.stack([LX.evaluate(p) for p in probes], -2),
               "lin_deriv": torch.stack([LX.derivative(p) for p in probes], -2),
               "cub_eval": torch.stack([CX.evaluate(p) for p in probes], -2),
               "cub_deriv": torch.stack([CX.derivative(p) for p in probes], -2),
               "lin_eval_vec": LX.evaluate(probes), "cub_deriv_vec": CX.derivative(probes)}
        ev.append(rec)
    out["evaluate"] = ev
    save("interpolation", out)


# ----------------------------------------------------------------------------------------------------------------
def _run_cdeint(X, func, z0, t, w, **kw):
    """forward + backward through the reference; returns outputs and grads of loss = sum(out * w)."""
    for p in func.parameters():
        p.grad = None
    z0 = z0.clone().requires_grad_(True)
    if hasattr(func, "nfe"):
        func.nfe = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = torchcde.cdeint(X, func, z0, t, **kw)
    loss = (out * w).sum()
    loss.backward()
    nfe = getattr(func, "nfe", None)
    # named_parameters() de-duplicates the shared middle Linear
    grads = {n: p.grad.clone() for n, p in func.named_parameters()}
    return {"out": out.detach().clone(), "grad_z0": z0.grad.clone(), "grads": grads, "nfe": nfe}


def golden_cdeint():
    g = torch.Generator().manual_seed(4321)
    torch.manual_seed(99)
    out = {}

    def make_field(kind, C, H, HH, n):
        if kind == "orig":
            f = _vf.OriginalVectorField(input_dim=C, hidden_dim=H, hidden_hidden_dim=HH, num_layers=n)
        else:
            f = _toy.CDEFunc(C, H, width=HH)
        return f

    cases = [
        # name, field, B, L, C(incl. time), H, HH, n, interp, method, online, adjoint, options, t-mode
        ("c1_toy_rect_rk4", "toy", 8, 3, 2, 32, 128, 0, "rectilinear", "rk4", True, False, {"step_size": 1}, "grid"),
        ("c2_lin_rk4_term", "orig", 6, 12, 4, 16, 16, 3, "linear", "rk4", False, False, {"step_size": 1}, "interval"),
        ("c2_rect_rk4_online", "orig", 6, 9, 4, 16, 24, 3, "rectilinear", "rk4", True, False, {"step_size": 1}, "grid"),
        ("c2_lin_euler", "orig", 5, 10, 3, 8, 8, 1, "linear", "euler", True, False, {"step_size": 1}, "grid"),
        ("c5_small_rect", "orig", 4, 6, 10, 16, 16, 3, "rectilinear", "rk4", True, False, {"step_size": 1}, "grid"),
        ("cub_rk4_halfstep_offgrid", "orig", 3, 8, 3, 8, 12, 2, "cubic", "rk4", True, False, {"step_size": 0.5}, "offgrid"),
        ("lin_rk4_adjoint", "orig", 3, 7, 3, 8, 8, 2, "linear", "rk4", True, True, {"step_size": 1}, "grid"),
        ("c3_cub_dopri5", "orig", 4, 10, 5, 8, 8, 3, "cubic", "dopri5", False, False, {"min_step": 0.5}, "interval"),
        ("c3_cub_dopri5_adjoint", "orig", 4, 10, 5, 8, 8, 3, "cubic", "dopri5", False, True, {"min_step": 0.5}, "interval"),
        ("cub_dopri5_free_online", "orig", 3, 6, 3, 6, 8, 2, "cubic", "dopri5", True, False, {}, "grid"),
        ("cub_dopri5_free_adjoint_online", "orig", 3, 6, 3, 6, 8, 2, "cubic", "dopri5", True, True, {}, "grid"),
    ]
    for (name, kind, B, L, C, H, HH, n, interp, method, online, adjoint, options, tmode) in cases:
        x = torch.randn(B, L, C, generator=g)
        x[..., 0] = torch.arange(L, dtype=torch.float32)  # time channel as get_data/common.py:178-184
        x[..., 1:] = x[..., 1:].cumsum(-2) * 0.3
        if interp == "rectilinear":
            nan_mask_(x, 0.3, g, keep_first=True, skip_channel=0)
            coeffs = torchcde.linear_interpolation_coeffs(x, rectilinear=0)
        elif interp == "linear":
            coeffs = torchcde.linear_interpolation_coeffs(x)
        else:
            coeffs = torchcde.natural_cubic_coeffs(x)
        X = torchcde.NaturalCubicSpline(coeffs) if interp == "cubic" else torchcde.LinearInterpolation(coeffs)
        func = make_field(kind, C, H, HH, n)
        z0 = torch.randn(B, H, generator=g) * 0.5
        if tmode == "grid":
            t = X.grid_points
        elif tmode == "interval":
            t = X.interval
        else:
            lo, hi = X.interval
            t = torch.cat([lo.view(1), (lo + (hi - lo) * torch.rand(5, generator=g)).sort().values, hi.view(1)])
        w = torch.randn(B, len(t), H, generator=g)
        kw = dict(adjoint=adjoint, method=method, options=dict(options), atol=1e-5, rtol=1e-3)
        if method == "dopri5" and not options:
            kw.update(atol=1e-6, rtol=1e-4)
        res = _run_cdeint(X, func, z0, t, w, **kw)
        out[name] = {"x": x, "coeffs": coeffs, "interp": interp, "field": kind,
                     "dims": {"B": B, "L": L, "C": C, "H": H, "HH": HH, "n": n},
                     "state_dict": {k: v.clone() for k, v in func.state_dict().items()},
                     "z0": z0, "t": t, "w": w, "kw": kw, **res}
        print(name, "out", tuple(res["out"].shape), "nfe", res["nfe"])
    save("cdeint", out)


if __name__ == "__main__":
    golden_interpolation()
    golden_cdeint()
