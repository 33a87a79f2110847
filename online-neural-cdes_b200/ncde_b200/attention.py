"""AttentionNeuralCDE (mirrors src/ncde/attention.py of the reference) on top of torchcde_b200.

    1. hidden encoding         dZ = f(Z) dX                      (a NeuralCDE over the data path, online outputs)
    2. attention weights       dA = f(A) dZ, run backwards over Z (a NeuralCDE whose control path is the hidden sequence),
                               softmax / sparsemax over time
    3. keep the hidden states whose weight exceeds 1 / length, pad the ragged result by repeating the last kept state
    4. a final NeuralCDE over the kept states, then a linear read-out.

Every CDE solve runs in libncde_b200; the hidden sequences become control paths through torchcde_b200's
linear_interpolation_coeffs / LinearInterpolation on the device (reference: attention.py:113).  The reference pads the ragged
kept states with the un-vendored `autots` (PadRaggedTensors + ForwardFill); here that is one gather on the device:
row s of the result lists the kept states of series s in order and then repeats the last one.
"""
import torch
from torch import nn

import torchcde_b200 as torchcde

from . import NeuralCDE


class Sparsemax(nn.Module):
    """Sparsemax (Martins & Astudillo 2016): Euclidean projection of the logits onto the probability simplex along `dim`
    (src/ncde/sparsemax.py).  Sort-based: tau = (sum of the k largest logits - 1) / k with k the largest index for which
    1 + k z_(k) > sum_{j<=k} z_(j); output = max(z - tau, 0).  Differentiable through autograd."""

    def __init__(self, dim=None):
        super().__init__()
        self.dim = -1 if dim is None else dim

    def forward(self, logits):
        z = logits.movedim(self.dim, -1)
        z = z - z.max(dim=-1, keepdim=True)[0]
        zs = torch.sort(z, dim=-1, descending=True)[0]
        k = torch.arange(1, z.size(-1) + 1, device=z.device, dtype=z.dtype)
        csum = zs.cumsum(-1)
        support = (1 + k * zs) > csum
        k_max = (support * k).max(dim=-1, keepdim=True)[0]
        tau = ((support * zs).sum(-1, keepdim=True) - 1) / k_max
        return torch.clamp(z - tau, min=0).movedim(-1, self.dim)


class FlipTensor(nn.Module):
    """Flip a tensor (or item `item_index` of a list of tensors) along `dim` (attention.py:143-162)."""

    def __init__(self, dim=-2, item_index=None):
        super().__init__()
        self.dim = dim
        self.item_index = item_index

    def forward(self, x):
        if self.dim is None:
            return x
        if self.item_index is None:
            return x.flip(dims=[self.dim])
        # The reference indexes whatever it is given (attention.py:153-157): item `item_index` of the [static, sequence] list on the
        # way in — and, on the way out of the attention CDE, ROW `item_index` of the (batch, length, 1) tensor, i.e. with static
        # features only that one series is flipped back.  Kept as is: a drop-in must reproduce the reference's outputs.
        x = x.clone() if isinstance(x, torch.Tensor) else list(x)
        x[self.item_index] = x[self.item_index].flip(dims=[self.dim])
        return x


def keep_and_pad(hidden_state, keep):
    """hidden_state (B, T, H), keep (B, T) bool -> (B, max kept, H): the kept states of every series in order, the rest of
    the row repeating its last kept state (= NaN padding of the ragged list + forward fill, attention.py:101-111).  One gather;
    gradients flow to the gathered states."""
    B, T, H = hidden_state.shape
    counts = keep.sum(1)
    n_max = int(counts.max())            # the one host synchronisation: the output shape depends on the data
    if int(counts.min()) == 0:
        raise ValueError("a series kept no hidden state (every attention weight <= 1 / length)")
    # position of the j-th kept state of each row: stable sort of the mask puts kept indices first, in time order
    order = torch.sort((~keep).to(torch.int8), dim=1, stable=True)[1]              # (B, T)
    j = torch.arange(n_max, device=keep.device).unsqueeze(0).expand(B, -1)
    src = order.gather(1, torch.minimum(j, (counts - 1).unsqueeze(1)))              # (B, n_max) time indices, last one repeated
    return hidden_state.gather(1, src.unsqueeze(-1).expand(-1, -1, H))


class AttentionNeuralCDE(nn.Module):
    """Constructor arguments and sub-module names follow src/ncde/attention.py:10-141, so a reference state_dict loads."""

    def __init__(self, input_dim, hidden_dim, output_dim, static_dim=None, adjoint=True, run_backwards=True, sparsemax=False,
                 precision=None):
        super().__init__()
        self.input_dim, self.hidden_dim, self.output_dim = input_dim, hidden_dim, output_dim
        self.static_dim = static_dim
        self.adjoint = adjoint
        self.run_backwards = run_backwards
        self.precision = precision
        self.encoder = self._create_ncde(input_dim, hidden_dim, hidden_dim, static_dim)
        activation = Sparsemax(dim=1) if sparsemax else nn.Softmax(dim=1)
        self.attention = nn.Sequential(self._create_flipper(), self._create_ncde(hidden_dim, hidden_dim, 1, static_dim),
                                       self._create_flipper(), activation)
        self.final = nn.Sequential(self._create_ncde(hidden_dim, hidden_dim, hidden_dim, static_dim, return_sequences=False))
        self.fc_output = nn.Linear(hidden_dim, output_dim)

    def _create_flipper(self):
        if self.run_backwards:
            return FlipTensor(dim=-2, item_index=1 if self.static_dim else None)
        return FlipTensor(dim=None)

    def _create_ncde(self, input_dim, hidden_dim, output_dim, static_dim, return_sequences=True):
        return NeuralCDE(input_dim, hidden_dim, output_dim, static_dim, use_initial=True, interpolation="linear",
                         adjoint=self.adjoint, num_layers=3, apply_final_linear=True, return_sequences=return_sequences,
                         return_filtered_rectilinear=False, precision=self.precision)

    def _with_static(self, x, hidden_state):
        return hidden_state if self.static_dim is None else [x[0], hidden_state]

    def reduce_hidden_state(self, x, hidden_state, attention_weights):
        keep = (attention_weights > 1 / hidden_state.size(1)).reshape(hidden_state.size(0), -1)
        reduced = keep_and_pad(hidden_state, keep)
        return self._with_static(x, torchcde.linear_interpolation_coeffs(reduced))

    def forward(self, x):
        hidden_state = self.encoder(x)
        attention_weights = self.attention(self._with_static(x, hidden_state))
        reduced = self.reduce_hidden_state(x, hidden_state, attention_weights)
        return self.fc_output(self.final(reduced))
