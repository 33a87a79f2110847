import sys, torch, copy
import os; R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [R, os.path.join(R, "tests"), os.path.join(R, "online-neural-cdes_b200")]
from oracle import cde_oracle as O
import torchcde_b200 as tc
from test_gpu_solve import _config_case
for name in ["cfg2_linear", "cfg2_rect", "cfg4", "cfg5", "odd_shapes"]:
    x, func, z0, interp, online = _config_case(name)
    if interp == "rectilinear": cref = O.linear_interpolation_coeffs(x.clone(), rectilinear=0)
    elif interp == "linear": cref = O.linear_interpolation_coeffs(x.clone())
    else: cref = O.natural_cubic_coeffs(x.clone())
    def oracle(dtype):
        f = copy.deepcopy(func).to(dtype)
        Xr = (O.CubicPath if interp == "cubic" else O.LinearPath)(cref.to(dtype))
        t = Xr.grid_points if online else Xr.interval
        g = torch.Generator().manual_seed(11)
        w = torch.randn(x.shape[0], len(t), z0.shape[1], generator=g).to(dtype)
        z0r = z0.clone().to(dtype).requires_grad_(True)
        o = O.cdeint(Xr, f, z0r, t, adjoint=False, method="rk4", options={"step_size": 1})
        (o * w).sum().backward()
        return o.detach().double(), z0r.grad.double(), w
    o64, g64, _ = oracle(torch.float64)
    o32, g32, w = oracle(torch.float32)
    fd = copy.deepcopy(func).cuda()
    X = (tc.NaturalCubicSpline if interp == "cubic" else tc.LinearInterpolation)(cref.cuda())
    z0d = z0.cuda().requires_grad_(True)
    out = tc.cdeint(X, fd, z0d, X.grid_points if online else X.interval, adjoint=False, method="rk4", options={"step_size": 1})
    (out * w.float().cuda()).sum().backward()
    gd = z0d.grad.double().cpu()
    scale = g64.abs().max()
    row_err_gpu = (gd - g64).abs().amax(1) / scale
    row_err_cpu = (g32 - g64).abs().amax(1) / scale
    print(name, "rows", x.shape[0], "GPU rows>1e-5:", int((row_err_gpu > 1e-5).sum()), "max %.2e median %.2e" % (row_err_gpu.max(), row_err_gpu.median()),
          "| CPU32 rows>1e-5:", int((row_err_cpu > 1e-5).sum()), "max %.2e median %.2e" % (row_err_cpu.max(), row_err_cpu.median()),
          "| out err gpu %.2e cpu %.2e" % (float((out.double().cpu()-o64).abs().max()/o64.abs().max()), float((o32-o64).abs().max()/o64.abs().max())))
