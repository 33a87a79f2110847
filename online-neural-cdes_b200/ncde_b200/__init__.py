"""NeuralCDE model wrapper (mirrors src/ncde/ncde.py of the reference) on top of torchcde_b200."""
import torch
from torch import nn

import torchcde_b200 as torchcde

from .vector_fields import (VECTOR_FIELDS, GRUGatedVectorField, MinimalGatedVectorField,  # noqa: F401
                            OriginalVectorField)

from .interpolation import SmoothLinearInterpolation  # noqa: E402

SPLINES = {
    "cubic": torchcde.NaturalCubicSpline,
    "linear": torchcde.LinearInterpolation,
    "rectilinear": torchcde.LinearInterpolation,
    "linear_cubic_smoothing": SmoothLinearInterpolation,
    "linear_quintic_smoothing": SmoothLinearInterpolation,
}


class NeuralCDE(nn.Module):
    """h0 = Linear([static ⊕] X(0)); hidden = cdeint(X, f_theta, h0, times); outputs = Linear(hidden).

    Constructor arguments, defaults, tolerances (atol 1e-5, rtol 1e-3) and solver options ({'step_size': 1} for
    rk4, {'min_step': 0.5} for dopri5) follow src/ncde/ncde.py:44-134.  ``precision`` selects the arithmetic of the
    final-layer tiles of the fused solve ('fp32' or 'bf16' tensor-core tiles).
    """

    def __init__(self, input_dim, hidden_dim, output_dim, static_dim=None, hidden_hidden_dim=15, num_layers=3,
                 use_initial=True, interpolation="linear", interpolation_eps=None, sparsity=None,
                 vector_field="original", vector_field_type="matmul", adjoint=True, solver="rk4",
                 return_sequences=False, apply_final_linear=True, return_filtered_rectilinear=True,
                 precision=None):
        super().__init__()
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim
        self.output_dim = output_dim
        self.static_dim = static_dim
        self.hidden_hidden_dim = hidden_hidden_dim
        self.num_layers = num_layers
        self.use_initial = use_initial
        self.interpolation = interpolation
        self.vector_field_type = vector_field_type
        self.adjoint = adjoint
        self.solver = solver
        self.return_sequences = return_sequences
        self.apply_final_linear = apply_final_linear
        self.return_filtered_rectilinear = return_filtered_rectilinear
        self.precision = precision

        if self.initial_dim > 0:
            self.initial_linear = nn.Linear(self.initial_dim, self.hidden_dim)
        assert self.interpolation in SPLINES.keys(), "Unrecognised interpolation scheme {}".format(self.interpolation)
        # src/ncde/ncde.py:112-126
        if interpolation in ("linear_cubic_smoothing", "linear_quintic_smoothing"):
            match_second = "quintic" in interpolation
            self.spline = lambda coeffs: SmoothLinearInterpolation(coeffs, gradient_matching_eps=interpolation_eps,
                                                                   match_second_derivatives=match_second)
        else:
            if interpolation_eps == 1:
                interpolation_eps = None
            assert interpolation_eps is None
            self.spline = SPLINES[self.interpolation]

        assert self.solver in ["rk4", "dopri5"]
        self.atol = 1e-5
        self.rtol = 1e-3
        self.cdeint_options = {"step_size": 1} if self.solver == "rk4" else {"min_step": 0.5}

        if vector_field not in VECTOR_FIELDS:
            raise NotImplementedError("vector field '{}' is not implemented".format(vector_field))
        self.func = VECTOR_FIELDS[vector_field](input_dim=input_dim, hidden_dim=hidden_dim,
                                                hidden_hidden_dim=hidden_hidden_dim, num_layers=num_layers,
                                                sparsity=sparsity, vector_field_type=vector_field_type)
        self.final_linear = nn.Linear(self.hidden_dim, self.output_dim) if apply_final_linear else (lambda x: x)

    @property
    def initial_dim(self):
        d = 0
        if self.use_initial:
            d += self.input_dim
        if self.static_dim is not None:
            d += self.static_dim
        return d

    @property
    def nfe(self):
        return getattr(self.func, "nfe", None)

    def _setup_h0(self, inputs):
        """src/ncde/ncde.py:170-198."""
        if not self.static_dim:
            spline = self.spline(inputs)
            if self.use_initial:
                h0 = self.initial_linear(spline.evaluate(0))
            else:
                h0 = torch.zeros(inputs.size(0), self.hidden_dim, device=inputs.device)
        else:
            assert len(inputs) == 2, "Inputs must be a 2-tuple of (static_data, temporal_data)"
            static, spline = inputs[0], self.spline(inputs[1])
            if self.use_initial:
                h0 = self.initial_linear(torch.cat((static, spline.evaluate(0)), dim=-1))
            else:
                h0 = self.initial_linear(static)
        return spline, h0

    def _make_outputs(self, hidden):
        """src/ncde/ncde.py:200-212."""
        if self.return_sequences:
            outputs = self.final_linear(hidden)
            if self.interpolation == "rectilinear" and self.return_filtered_rectilinear:
                outputs = outputs[:, ::2]
        else:
            outputs = self.final_linear(hidden[:, -1, :])
        return outputs

    def forward(self, inputs):
        spline, h0 = self._setup_h0(inputs)
        times = spline.grid_points if self.return_sequences else spline.interval
        options = dict(self.cdeint_options)
        if self.precision is not None:
            options["precision"] = self.precision
        hidden = torchcde.cdeint(spline, self.func, h0, t=times, adjoint=self.adjoint,
                                 vector_field_type=self.vector_field_type, method=self.solver, atol=self.atol,
                                 rtol=self.rtol, options=options)
        return self._make_outputs(hidden)


from .stacked import StackedNeuralCDE  # noqa: E402,F401
from .attention import AttentionNeuralCDE, Sparsemax  # noqa: E402,F401
