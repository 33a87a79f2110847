"""torchcde_b200 — B200-native drop-in for the ``torchcde`` solve path used by jambo6/online-neural-cdes.

    import torchcde_b200 as torchcde
    coeffs = torchcde.linear_interpolation_coeffs(x, rectilinear=0)
    X = torchcde.LinearInterpolation(coeffs)
    z = torchcde.cdeint(X, func, z0, X.grid_points, adjoint=False, method='rk4', options={'step_size': 1})

Same names as modules/torchcde/torchcde/__init__.py:1-6 of the reference.  Everything numerical runs in
libncde_b200.so (hand-written sm_100a CUDA behind the C ABI of include/ncde_b200.h); there is no CPU fallback.
"""
from .interpolation_base import InterpolationBase
from .interpolation_cubic import (natural_cubic_spline_coeffs, natural_cubic_coeffs, NaturalCubicSpline,
                                  CubicSpline)
from .interpolation_linear import linear_interpolation_coeffs, LinearInterpolation
from .log_ode import logsig_windows, logsignature_windows, logsignature_channels
from .misc import TupleControl, forward_fill
from .solver import cdeint
from . import distributed  # noqa: F401

__version__ = "0.1.0"
