import abc

import torch

from . import _capi


def path_eval_raw(kind, coeffs, derivs, knots, tq, deriv, C, index=None):
    """ncde_path_eval on contiguous, detached tensors -> (..., n_t, C)."""
    K = coeffs.size(-2) + (1 if kind == _capi.PATH_CUBIC else 0)
    per = coeffs.size(-2) * coeffs.size(-1)
    n = coeffs.numel() // per if coeffs.numel() else 0
    out = torch.empty(*coeffs.shape[:-2], tq.numel(), C, dtype=coeffs.dtype, device=coeffs.device)
    _capi.check(_capi.lib().ncde_path_eval(kind, _capi.dtype_code(coeffs), coeffs.data_ptr(), _capi.ptr(derivs),
                                           knots.data_ptr(), n, K, C, tq.data_ptr(), tq.numel(), int(deriv),
                                           out.data_ptr(), _capi.ptr(index), _capi.stream_ptr(coeffs.device)))
    return out


class PathEvalGrad(torch.autograd.Function):
    """``evaluate`` / ``derivative`` for coefficient tensors that require gradients (stacked Neural CDEs:
    h0 = Linear(X.evaluate(0)), src/ncde/ncde.py:179-181).  Forward is ncde_path_eval, backward ncde_path_eval_bwd — no
    torch arithmetic on either side."""

    @staticmethod
    def forward(ctx, coeffs, kind, derivs, knots, tq, deriv, C):
        ctx.kind, ctx.deriv, ctx.C = kind, deriv, C
        ctx.shape, ctx.dtype = coeffs.shape, coeffs.dtype
        ctx.save_for_backward(knots, tq)
        return path_eval_raw(kind, coeffs.detach().contiguous(), derivs, knots, tq, deriv, C)

    @staticmethod
    def backward(ctx, grad_out):
        knots, tq = ctx.saved_tensors
        C = ctx.C
        g = grad_out.contiguous()
        grad = torch.zeros(ctx.shape, dtype=ctx.dtype, device=g.device)
        K = ctx.shape[-2] + (1 if ctx.kind == _capi.PATH_CUBIC else 0)
        n = g.numel() // (tq.numel() * C) if g.numel() else 0
        _capi.check(_capi.lib().ncde_path_eval_bwd(ctx.kind, _capi.dtype_code(grad), knots.data_ptr(), n, K, C,
                                                   tq.data_ptr(), tq.numel(), int(ctx.deriv), g.data_ptr(),
                                                   grad.data_ptr(), _capi.stream_ptr(g.device)))
        return grad, None, None, None, None, None, None


class InterpolationBase(torch.nn.Module, metaclass=abc.ABCMeta):
    """Same abstract interface as torchcde/interpolation_base.py:5-22."""

    @property
    @abc.abstractmethod
    def grid_points(self):
        """The knots of the interpolation (one dimensional tensor)."""

    @property
    @abc.abstractmethod
    def interval(self):
        """Two element tensor: first and last knot."""

    @abc.abstractmethod
    def evaluate(self, t):
        """Value of the path at t (any shape) -> (..., *t.shape, channels)."""

    @abc.abstractmethod
    def derivative(self, t):
        """Derivative of the path at t (any shape) -> (..., *t.shape, channels)."""
