"""StackedNeuralCDE (mirrors src/ncde/stacked.py of the reference): a chain of Neural CDEs in which the hidden path of
one is the control path of the next,

    dZ_1 = f_1(Z_1) dX,   dZ_2 = f_2(Z_2) dZ_1,   ...,   Y = L(Z_n).

Every link is a ``torchcde_b200.cdeint`` solve; with ``adjoint=False`` the gradient flows back through the intermediate
paths via ``ncde_solve_bwd``'s ``grad_coeffs`` output and ``ncde_path_eval_bwd`` (h0 = Linear(X.evaluate(0))).
"""
from torch import nn

from . import NeuralCDE


class StackedNeuralCDE(nn.Module):
    """Same constructor arguments and forward contract as src/ncde/stacked.py:7-131.  ``precision`` is forwarded to every
    link; gradients through the intermediate paths need 'fp32' (the default)."""

    def __init__(self, input_dim, hidden_dims, output_dim, hidden_hidden_dim=15, static_dim=None, adjoint=True,
                 return_sequences=False, static_in_all_layers=False, precision=None):
        assert isinstance(hidden_dims, list), "hidden_dims must be a list, got type {}".format(type(hidden_dims))
        super(StackedNeuralCDE, self).__init__()
        self.input_dim = input_dim
        self.hidden_dims = hidden_dims
        self.output_dim = output_dim
        self.hidden_hidden_dim = hidden_hidden_dim
        self.static_dim = static_dim
        self.adjoint = adjoint
        self.return_sequences = return_sequences
        self.static_in_all_layers = static_in_all_layers
        self.num_stacked = len(hidden_dims)

        # stacked.py:66-87: the final linear map and the requested output format apply to the last link only
        input_, static_ = input_dim, static_dim
        self.ncdes = nn.ModuleList()
        for i, hidden_ in enumerate(hidden_dims):
            last = i == self.num_stacked - 1
            # like the reference (stacked.py:101-121) `hidden_hidden_dim` is NOT forwarded: every link keeps NeuralCDE's
            # default of 15, so state_dicts are interchangeable
            self.ncdes.append(NeuralCDE(input_, hidden_, output_dim, static_,
                                        use_initial=True, interpolation="linear", adjoint=adjoint, num_layers=3,
                                        apply_final_linear=last,
                                        return_sequences=(self.return_sequences if last else True),
                                        precision=precision))
            input_ = hidden_
            if not self.static_in_all_layers:
                static_ = None
        # stacked.py:90 (defined but not applied by the reference's forward either; kept for state_dict compatibility)
        self.fc_output = nn.Linear(hidden_dims[-1], output_dim)

    def _handle_hidden_static_features(self, x, hidden_state):
        if any([self.static_dim is None, not self.static_in_all_layers]):
            return hidden_state
        return [x[0], hidden_state]

    def forward(self, x):
        hidden_state = self.ncdes[0](x)
        for ncde in self.ncdes[1:]:
            hidden_state = ncde(self._handle_hidden_static_features(x, hidden_state))
        return hidden_state
