"""Vector fields of the Neural CDE model (mirrors src/ncde/vector_fields/base.py of the reference).

The modules only hold parameters and define the architecture; inside ``torchcde_b200.cdeint`` they are lowered to an
MLP descriptor and evaluated by the fused CUDA kernels, never called per stage.
"""
import torch
from torch import nn


class OriginalVectorField(nn.Module):
    """f_theta: R^H -> R^{H x C}:  Linear(H, HH) + ReLU, (num_layers - 1) x [Linear(HH, HH) + ReLU], Linear(HH, H*C) +
    tanh, view(-1, H, C).  With vector_field_type 'evaluate' / 'derivative' (base.py:56-60): R^{H+C} -> R^H, no view.

    Like the reference (base.py:64-69) the middle layers are ONE Linear module repeated, so they share weights and
    their gradient accumulates over the repeats.  ``nfe`` counts vector-field evaluations (base.py:61,90).
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
        self.tanh_output_layer = nn.Sequential(nn.Linear(hidden_hidden_dim, self.output_dim), nn.Tanh())

    def forward(self, t, h):
        """Eager definition (used once to validate the lowering; the solve never calls it)."""
        out = self.tanh_output_layer(self.net_to_hh(h))
        if self.matmul:
            out = out.view(-1, self.hidden_dim, self.input_dim)
        self.nfe += 1
        return out


VECTOR_FIELDS = {"original": OriginalVectorField}
