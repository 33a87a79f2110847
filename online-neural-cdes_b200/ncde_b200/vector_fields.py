"""Vector fields of the Neural CDE model (mirrors src/ncde/vector_fields/base.py and gating.py of the reference).

The modules only hold parameters and define the architecture; inside ``torchcde_b200.cdeint`` they are lowered to an
MLP descriptor and evaluated by the fused CUDA kernels, never called per stage.
"""
import torch
from torch import nn


class BaseVectorField(nn.Module):
    """Common part (base.py:7-92): ``net_to_hh`` = Linear(initial_dim, HH) + ReLU, (num_layers - 1) x [Linear(HH, HH) + ReLU].

    Like the reference (base.py:64-69) the middle layers are ONE Linear module repeated, so they share weights and
    their gradient accumulates over the repeats.  ``nfe`` counts vector-field evaluations (base.py:61,90).
    With vector_field_type 'evaluate' / 'derivative' (base.py:56-60) the field maps R^{H+C} -> R^H and is not reshaped.
    """

    def __init__(self, input_dim, hidden_dim, hidden_hidden_dim=15, num_layers=1, sparsity=None,
                 vector_field_type="matmul"):
        super().__init__()
        if vector_field_type not in ("matmul", "evaluate", "derivative"):
            raise ValueError("vector_field_type string not recognised")
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim
        self.hidden_hidden_dim = hidden_hidden_dim
        self.num_layers = num_layers
        self.sparsity = sparsity
        self.vector_field_type = vector_field_type
        self.nfe = 0
        # base.py:56-60: the control (or its derivative) is concatenated to the state unless the field is contracted with dX/dt
        self.matmul = vector_field_type == "matmul"
        self.initial_dim = hidden_dim if self.matmul else hidden_dim + input_dim
        self.output_dim = hidden_dim * input_dim if self.matmul else hidden_dim
        first = nn.Linear(self.initial_dim, hidden_hidden_dim)
        mods = [first, nn.ReLU()]
        if num_layers > 1:
            shared = nn.Linear(hidden_hidden_dim, hidden_hidden_dim)
            for _ in range(num_layers - 1):
                mods += [shared, nn.ReLU()]
        self.net_to_hh = nn.Sequential(*mods)
        self.additional_network_initialisation()

    def additional_network_initialisation(self):
        raise NotImplementedError

    def _forward(self, h):
        raise NotImplementedError

    def forward(self, t, h):
        """Eager definition (used once to validate the lowering; the solve never calls it)."""
        out = self._forward(h)
        if self.matmul:
            out = out.view(-1, self.hidden_dim, self.input_dim)
        self.nfe += 1
        return out


class OriginalVectorField(BaseVectorField):
    """f_theta: net_to_hh, then Linear(HH, H*C) + tanh (base.py:94-104)."""

    def additional_network_initialisation(self):
        self.tanh_output_layer = nn.Sequential(nn.Linear(self.hidden_hidden_dim, self.output_dim), nn.Tanh())

    def _forward(self, h):
        return self.tanh_output_layer(self.net_to_hh(h))


class MinimalGatedVectorField(BaseVectorField):
    """sigmoid(Linear_z(hh)) * tanh(Linear_r(hh)) with hh = net_to_hh(h) (gating.py:7-32).  Lowered to a gated last layer
    (ncde_mlp_t.W_gate): both heads are one interleaved GEMM inside the final-layer kernels."""

    def additional_network_initialisation(self):
        assert self.sparsity is None, "sparsity not implemented for gated methods"
        self.sigmoid_net = nn.Sequential(nn.Linear(self.hidden_hidden_dim, self.output_dim), nn.Sigmoid())
        self.tanh_net = nn.Sequential(nn.Linear(self.hidden_hidden_dim, self.output_dim), nn.Tanh())

    def _forward(self, h):
        hh = self.net_to_hh(h)
        return self.sigmoid_net(hh) * self.tanh_net(hh)


class GRUGatedVectorField(BaseVectorField):
    """GRU-style gating (gating.py:35-61): sigmoid_net(net(h)) * tanh_net(net(reset_net(h) * h)).  Lowered to one widened chain
    that carries both evaluations of ``net_to_hh`` side by side (torchcde_b200.lowering._lower_gru); needs
    2 * hidden_hidden_dim <= 256."""

    def additional_network_initialisation(self):
        assert self.sparsity is None, "sparsity not implemented for gated methods"
        self.reset_net = nn.Sequential(nn.Linear(self.initial_dim, self.initial_dim), nn.Sigmoid())
        self.sigmoid_net = nn.Sequential(nn.Linear(self.hidden_hidden_dim, self.output_dim), nn.Sigmoid())
        self.tanh_net = nn.Sequential(nn.Linear(self.hidden_hidden_dim, self.output_dim), nn.Tanh())

    def _forward(self, h):
        inner = self.net_to_hh(h)
        reset = self.net_to_hh(self.reset_net(h) * h)
        return self.sigmoid_net(inner) * self.tanh_net(reset)


# src/ncde/ncde.py:23-29 ('sparse' / 'low-rank' are commented out there too: they need the un-vendored `sparselinear`)
VECTOR_FIELDS = {"original": OriginalVectorField, "minimal": MinimalGatedVectorField, "gru": GRUGatedVectorField}
