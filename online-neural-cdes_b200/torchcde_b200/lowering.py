"""Lower a vector-field ``func`` (an nn.Module mapping z:(B,H) -> (B,H,C)) to the MLP descriptor the CUDA kernels run.

The fused solve cannot call back into Python per stage, so ``func`` must be a chain of Linear layers with ReLU /
tanh / identity activations ending in ``tanh`` and a ``view(-1, H, C)`` — which is what every vector field on the
reference's hot path is (src/ncde/vector_fields/base.py:64-104 ``OriginalVectorField``,
experiments/sim_bm_toy_example.py:10-30 ``CDEFunc``, modules/torchcde/example/example.py:20-52).  Anything else raises:
there is no eager fallback.

Recognition order:
  1. ``func.ncde_mlp_spec()`` if the module provides it -> list of (weight, bias, activation) tuples;
  2. the ``net_to_hh`` + ``tanh_output_layer`` structure of the reference's vector fields;
  3. a torch.fx trace of ``func.forward(t, z)`` that is a straight Linear/activation chain.
Layers that reuse the same nn.Linear object (the reference repeats one Linear for all middle layers, base.py:65-68)
share a slot, so their weight gradient accumulates into one tensor.
"""
import torch
import torch.fx

from . import _capi

_ACT = {"none": _capi.ACT_NONE, "relu": _capi.ACT_RELU, "tanh": _capi.ACT_TANH}


class MlpSpec:
    """layers: list of (weight, bias_or_None, activation_code); slot[i] identifies shared parameters."""

    def __init__(self, layers, gate=None):
        # gate: (weight, bias_or_None) of a sigmoid head multiplying the (tanh) last layer — MinimalGatedVectorField
        self.gate = gate
        if gate is not None and len(layers) >= _capi.MAX_LAYERS:
            raise NotImplementedError("gated vector fields take at most {} Linear layers".format(_capi.MAX_LAYERS - 1))
        if not 1 <= len(layers) <= _capi.MAX_LAYERS:
            raise NotImplementedError("vector fields with {} Linear layers are not supported (max {})".format(
                len(layers), _capi.MAX_LAYERS))
        self.weights = [l[0] for l in layers]
        self.biases = [l[1] for l in layers]
        self.acts = [l[2] for l in layers]
        seen = {}
        self.slots = []
        for w in self.weights:
            self.slots.append(seen.setdefault(id(w), len(seen)))
        for i in range(1, len(layers)):
            if self.weights[i].shape[1] != self.weights[i - 1].shape[0]:
                raise ValueError("vector field layer {} takes {} inputs but the previous layer produces {}".format(
                    i, self.weights[i].shape[1], self.weights[i - 1].shape[0]))

    @property
    def unique_params(self):
        """Parameters in first-use order without duplicates: [(tensor, kind, layer_index)]"""
        out, seen = [], set()
        for i, (w, b) in enumerate(zip(self.weights, self.biases)):
            if id(w) not in seen:
                seen.add(id(w))
                out.append((w, "W", i))
                if b is not None:
                    out.append((b, "b", i))
        if self.gate is not None:   # the gate's gradients are returned in slot n_layers (ncde_b200.h, ncde_mlp_t.W_gate)
            out.append((self.gate[0], "W", len(self.weights)))
            if self.gate[1] is not None:
                out.append((self.gate[1], "b", len(self.weights)))
        return out

    def reference_forward(self, z, hidden, channels):
        """The same map written with torch ops; used only to validate a lowering once, never on the solve path.
        ``channels`` None: vector_field_type evaluate / derivative (the output is the (B, H) state derivative itself)."""
        last = len(self.weights) - 1
        for i, (w, b, a) in enumerate(zip(self.weights, self.biases, self.acts)):
            zin = z
            z = torch.nn.functional.linear(z, w, b)
            if a == _capi.ACT_RELU:
                z = z.relu()
            elif a == _capi.ACT_TANH:
                z = z.tanh()
            elif a == _capi.ACT_GATE_IN:
                d = zin.shape[-1]
                z = torch.cat([z[..., :d], z[..., d:].sigmoid() * zin], -1)
            if i == last and self.gate is not None:
                z = torch.nn.functional.linear(zin, self.gate[0], self.gate[1]).sigmoid() * z
        return z.view(-1, hidden) if channels is None else z.view(-1, hidden, channels)


def _from_sequential(mods):
    layers = []
    for m in mods:
        if isinstance(m, torch.nn.Linear):
            layers.append([m.weight, m.bias, _capi.ACT_NONE])
        elif isinstance(m, torch.nn.ReLU):
            layers[-1][2] = _capi.ACT_RELU
        elif isinstance(m, torch.nn.Tanh):
            layers[-1][2] = _capi.ACT_TANH
        elif isinstance(m, torch.nn.Identity):
            pass
        else:
            raise NotImplementedError("unsupported module {} in vector field".format(type(m).__name__))
    return layers


def _from_fx(func):
    try:
        gm = torch.fx.symbolic_trace(func)
    except Exception as e:  # noqa: BLE001
        raise NotImplementedError("could not trace the vector field ({}); provide func.ncde_mlp_spec()".format(e))
    layers = []
    placeholders = [n for n in gm.graph.nodes if n.op == "placeholder"]
    if len(placeholders) < 2:
        raise NotImplementedError("vector field must have signature forward(t, z)")
    cur = placeholders[1]
    mods = dict(gm.named_modules())

    def set_act(code):
        if not layers or layers[-1][2] != _capi.ACT_NONE:
            raise NotImplementedError("activation without a preceding Linear layer")
        layers[-1][2] = code

    for node in gm.graph.nodes:
        if node.op in ("placeholder", "get_attr"):
            continue
        if node.op == "output":
            break
        inputs = [a for a in node.args if isinstance(a, torch.fx.Node)]
        tname = getattr(node.target, "__name__", str(node.target))
        # size()/getitem nodes used to build the final view are ignored
        if node.op == "call_method" and node.target in ("size", "dim"):
            continue
        if node.op == "call_function" and tname in ("getitem", "mul", "add", "neg", "floordiv"):
            if cur not in inputs:
                continue
        if cur not in inputs:
            raise NotImplementedError("vector field is not a straight chain (node {})".format(node.name))
        if node.op == "call_module":
            m = mods[node.target]
            if isinstance(m, torch.nn.Linear):
                layers.append([m.weight, m.bias, _capi.ACT_NONE])
            elif isinstance(m, torch.nn.ReLU):
                set_act(_capi.ACT_RELU)
            elif isinstance(m, torch.nn.Tanh):
                set_act(_capi.ACT_TANH)
            elif isinstance(m, torch.nn.Identity):
                pass
            else:
                raise NotImplementedError("unsupported module {} in vector field".format(type(m).__name__))
        elif node.op == "call_method" and node.target in ("relu", "tanh"):
            set_act(_ACT[node.target])
        elif node.op == "call_function" and tname in ("relu", "tanh"):
            set_act(_ACT[tname])
        elif (node.op == "call_method" and node.target in ("view", "reshape", "contiguous")) or \
                (node.op == "call_function" and tname in ("reshape",)):
            pass
        else:
            raise NotImplementedError("unsupported operation {} {} in vector field".format(node.op, node.target))
        cur = node
    return layers


_CACHE = {}
_GRU_VALIDATED = {}


def _lower_gru(func, hidden, channels, vector_field_type):
    """GRUGatedVectorField (src/ncde/vector_fields/gating.py:35-61):

        out = sigmoid(W_z net(x) + b_z) * tanh(W_r net(sigmoid(W_f x + b_f) * x) + b_r),   net = net_to_hh (shared weights)

    lowered to ONE chain that carries both evaluations of ``net`` side by side:
        layer 0   W = [I ; W_f], gate-in activation       x -> [x ; sigmoid(W_f x + b_f) * x]
        layer k   W = blockdiag(W_k, W_k), ReLU            [a ; a'] -> [relu(W_k a + b_k) ; relu(W_k a' + b_k)]
        last      tanh head [0 | W_r] gated by the sigmoid head [W_z | 0]
    The widened matrices are assembled here with differentiable concatenations of the module's parameters, on every call (the
    parameters change every optimiser step), so autograd folds the gradients the library returns for them back into W_f, W_k
    (both diagonal blocks, every repeat of the shared Linear), W_z and W_r; the blocks of zeros / the identity receive gradients
    that are discarded.  Needs 2 * hidden_hidden_dim <= 256 (the final-layer kernels' input width)."""
    if getattr(func, "vector_field_type", vector_field_type) != vector_field_type:
        raise ValueError("the vector field was built for vector_field_type='{}' but cdeint was called with '{}'".format(
            func.vector_field_type, vector_field_type))
    matmul = vector_field_type == "matmul"
    mods = list(func.net_to_hh)
    reset, sig, tanh = list(func.reset_net), list(func.sigmoid_net), list(func.tanh_net)
    ok = len(reset) == 2 and isinstance(reset[0], torch.nn.Linear) and isinstance(reset[1], torch.nn.Sigmoid) and \
        len(sig) == 2 and isinstance(sig[0], torch.nn.Linear) and isinstance(sig[1], torch.nn.Sigmoid) and \
        len(tanh) == 2 and isinstance(tanh[0], torch.nn.Linear) and isinstance(tanh[1], torch.nn.Tanh) and \
        len(mods) % 2 == 0 and all(isinstance(m, torch.nn.Linear) for m in mods[0::2]) and \
        all(isinstance(m, torch.nn.ReLU) for m in mods[1::2])
    if not ok:
        raise NotImplementedError("unexpected structure of the GRU-gated vector field")
    wf, bf = reset[0].weight, reset[0].bias
    d0 = wf.shape[1]
    zeros = wf.new_zeros
    eye = torch.eye(d0, dtype=wf.dtype, device=wf.device)
    layers = [(torch.cat([eye, wf], 0), torch.cat([zeros(d0), bf if bf is not None else zeros(d0)]), _capi.ACT_GATE_IN)]
    for lin in mods[0::2]:
        b = lin.bias if lin.bias is not None else zeros(lin.out_features)
        layers.append((torch.block_diag(lin.weight, lin.weight), torch.cat([b, b]), _capi.ACT_RELU))
    wz, wr = sig[0].weight, tanh[0].weight
    pad = torch.zeros_like(wz)
    gate = (torch.cat([wz, pad], 1), sig[0].bias)
    layers.append((torch.cat([pad, wr], 1), tanh[0].bias, _capi.ACT_TANH))
    spec = MlpSpec(layers, gate)
    spec.vector_field_type = vector_field_type
    spec.channels = channels
    d_in = hidden if matmul else hidden + channels
    d_out = hidden * channels if matmul else hidden
    if d0 != d_in or wr.shape[0] != d_out:
        raise ValueError("vector field maps {} -> {} but the solve (vector_field_type='{}') needs {} -> {}".format(
            d0, wr.shape[0], vector_field_type, d_in, d_out))
    hit = _GRU_VALIDATED.get(id(func))
    if hit is None or hit() is not func:
        nfe_before = getattr(func, "nfe", None)
        with torch.no_grad():
            probe = torch.linspace(-1.0, 1.0, 3 * d_in, dtype=wf.dtype, device=wf.device).view(3, d_in)
            want = func(torch.zeros((), dtype=wf.dtype, device=wf.device), probe)
            got = spec.reference_forward(probe, hidden, channels if matmul else None)
            if want.shape != got.shape or not torch.allclose(want, got, rtol=1e-4, atol=1e-5):
                raise NotImplementedError("GRU-gated vector field could not be lowered faithfully")
        if nfe_before is not None:
            func.nfe = nfe_before
        import weakref
        _GRU_VALIDATED[id(func)] = weakref.ref(func)
    return spec


def lower(func, hidden, channels, vector_field_type="matmul"):
    """Return the MlpSpec of ``func``; validated once per module object against func itself.  For vector_field_type
    'evaluate' / 'derivative' (torchcde/solver.py:123-126) func maps [z, X(t) | dX/dt(t)]: (B, H+C) -> (B, H)."""
    key = id(func)
    matmul = vector_field_type == "matmul"
    hit = _CACHE.get(key)
    if hit is not None and hit[0]() is func and hit[2] == (hidden, channels, vector_field_type):
        return hit[1]
    nfe_before = getattr(func, "nfe", None)
    gate = None
    if hasattr(func, "reset_net"):
        return _lower_gru(func, hidden, channels, vector_field_type)
    if hasattr(func, "ncde_mlp_spec"):
        layers = [[w, b, _ACT[a] if isinstance(a, str) else a] for (w, b, a) in func.ncde_mlp_spec()]
    elif isinstance(getattr(func, "net_to_hh", None), torch.nn.Sequential) and \
            isinstance(getattr(func, "tanh_output_layer", None), torch.nn.Sequential):
        if getattr(func, "vector_field_type", vector_field_type) != vector_field_type:
            raise ValueError("the vector field was built for vector_field_type='{}' but cdeint was called with '{}'".format(
                func.vector_field_type, vector_field_type))
        layers = _from_sequential(list(func.net_to_hh) + list(func.tanh_output_layer))
    elif isinstance(getattr(func, "net_to_hh", None), torch.nn.Sequential) and \
            isinstance(getattr(func, "sigmoid_net", None), torch.nn.Sequential) and \
            isinstance(getattr(func, "tanh_net", None), torch.nn.Sequential):
        # MinimalGatedVectorField (gating.py:7-32): sigmoid_net(hh) * tanh_net(hh)
        if getattr(func, "vector_field_type", vector_field_type) != vector_field_type:
            raise ValueError("the vector field was built for vector_field_type='{}' but cdeint was called with '{}'".format(
                func.vector_field_type, vector_field_type))
        layers = _from_sequential(list(func.net_to_hh) + list(func.tanh_net))
        sig = list(func.sigmoid_net)
        if len(sig) != 2 or not isinstance(sig[0], torch.nn.Linear) or not isinstance(sig[1], torch.nn.Sigmoid):
            raise NotImplementedError("sigmoid_net must be Linear + Sigmoid")
        gate = (sig[0].weight, sig[0].bias)
    elif isinstance(func, torch.nn.Module):
        layers = _from_fx(func)
    else:
        raise NotImplementedError("func must be an nn.Module that is a Linear/activation chain (no eager fallback)")
    if not layers:
        raise NotImplementedError("vector field has no Linear layer")
    spec = MlpSpec([tuple(l) for l in layers], gate)
    if gate is not None and gate[0].shape != spec.weights[-1].shape:
        raise ValueError("the gate of the vector field must have the shape of its last layer")
    spec.vector_field_type = vector_field_type
    spec.channels = channels
    d_in = hidden if matmul else hidden + channels
    d_out = hidden * channels if matmul else hidden
    if spec.weights[0].shape[1] != d_in or spec.weights[-1].shape[0] != d_out:
        raise ValueError("vector field maps {} -> {} but the solve (vector_field_type='{}') needs {} -> {}".format(
            spec.weights[0].shape[1], spec.weights[-1].shape[0], vector_field_type, d_in, d_out))
    if spec.acts[-1] != _capi.ACT_TANH:
        raise NotImplementedError("the vector field must end in tanh (as every vector field of the reference does)")
    # one-off structural validation on three probe rows
    with torch.no_grad():
        w0 = spec.weights[0]
        probe = torch.linspace(-1.0, 1.0, 3 * d_in, dtype=w0.dtype, device=w0.device).view(3, d_in)
        want = func(torch.zeros((), dtype=w0.dtype, device=w0.device), probe)
        got = spec.reference_forward(probe, hidden, channels if matmul else None)
        if want.shape != got.shape or not torch.allclose(want, got, rtol=1e-4, atol=1e-5):
            raise NotImplementedError("vector field could not be lowered to a Linear/activation chain faithfully")
    if nfe_before is not None:
        func.nfe = nfe_before
    import weakref
    try:
        ref = weakref.ref(func)
    except TypeError:
        ref = (lambda f: (lambda: f))(func)
    _CACHE[key] = (ref, spec, (hidden, channels, vector_field_type))
    return spec
