import abc

import torch


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
