"""dopri5 (adaptive) solve — device-side step control.  Not built yet in this revision."""


def solve(X, func, spec, z0, t, t_host, adjoint, options, kwargs):
    raise NotImplementedError("method='dopri5' is not implemented yet in torchcde_b200")
