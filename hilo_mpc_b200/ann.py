"""Feed-forward network inference graph (the part of the reference's ANN support that lies on the hot path).

The reference trains MLPs with its PyTorch/TensorFlow plugins and then exports the trained weights to a CasADi graph so
that the network can be embedded in a ``Model`` and differentiated by IPOPT (``net_to_casadi_graph``,
``hilo_mpc/util/machine_learning.py:521-578``; activations :44-118; ``ANN.predict`` ``nn/nn.py:536-544``;
``Model.substitute_from`` ``dynamic_model.py:3040-3125``).  Training is out of scope (PyTorch already does it on the
GPU); this class takes trained weights (e.g. from a ``torch.nn.Sequential``) and provides the same graph as
``hilo_mpc_b200.symbolic`` expressions, from which the code generator emits the forward pass and its Jacobian as
device code inside the node kernels.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from . import symbolic as sx

_ACT = {
    'sigmoid': sx.sigmoid, 'logistic': sx.sigmoid, 'tanh': sx.tanh, 'linear': lambda v: v, 'identity': lambda v: v,
    'softplus': lambda v: sx.log(1. + sx.exp(v)), 'soft_plus': lambda v: sx.log(1. + sx.exp(v)),
}


class Layer:
    def __init__(self, nodes: int, activation: str = 'sigmoid', type: str = 'dense'):
        self.nodes, self.activation, self.type = int(nodes), activation, type

    @staticmethod
    def dense(nodes, activation='sigmoid', **kwargs):
        return Layer(nodes, activation, 'dense')


class ANN:
    """``ANN(features, labels)`` + ``add_layers(Layer.dense(n, activation=...))`` + ``set_weights(...)`` ->
    ``predict(x)`` numeric (NumPy) or symbolic."""

    def __init__(self, features, labels, id=None, name=None, **kwargs):
        self.features = [features] if isinstance(features, str) else list(features)
        self.labels = [labels] if isinstance(labels, str) else list(labels)
        self.name = name or 'ann'
        self._layers: List[Layer] = []
        self._weights: Optional[List[np.ndarray]] = None
        self._bias: Optional[List[np.ndarray]] = None
        self._in_mean = self._in_scale = self._out_mean = self._out_scale = None

    symbolic_mode = 'call'        # 'call': single-hidden-layer networks as call nodes; 'unroll': plain expression graph
    call_node_min_hidden = 16     # smaller networks are cheaper unrolled

    def add_layers(self, *layers):
        for layer in layers:
            self._layers.extend(layer if isinstance(layer, (list, tuple)) else [layer])

    def setup(self, **kwargs):
        for layer in self._layers:
            if layer.activation.lower() not in _ACT:
                raise NotImplementedError(f"Activation '{layer.activation}' is not differentiable in the B200 "
                                          f"symbolic layer (supported: {sorted(_ACT)})")

    def set_weights(self, weights: Sequence, bias: Sequence, input_scaling=None, output_scaling=None):
        """``weights[k]`` has shape ``(n_out_k, n_in_k)`` (PyTorch ``Linear.weight`` layout), last entry = output layer.
        ``input_scaling`` / ``output_scaling``: ``(mean, scale)`` z-score pairs (machine_learning.py:543-565)."""
        self._weights = [np.asarray(w, dtype=float) for w in weights]
        self._bias = [np.asarray(b, dtype=float).reshape(-1) for b in bias]
        if len(self._weights) != len(self._layers) + 1:
            raise ValueError(f"Expected {len(self._layers) + 1} weight matrices (hidden layers + linear output layer), "
                             f"got {len(self._weights)}")
        if input_scaling is not None:
            self._in_mean, self._in_scale = (np.asarray(v, dtype=float).reshape(-1) for v in input_scaling)
        if output_scaling is not None:
            self._out_mean, self._out_scale = (np.asarray(v, dtype=float).reshape(-1) for v in output_scaling)

    @classmethod
    def from_torch(cls, features, labels, module, input_scaling=None, output_scaling=None, name=None):
        """Build from a ``torch.nn.Sequential`` of ``Linear`` + activation modules."""
        import torch.nn as nn
        acts = {nn.Tanh: 'tanh', nn.Sigmoid: 'sigmoid', nn.Softplus: 'softplus', nn.Identity: 'linear'}
        net = cls(features, labels, name=name)
        ws, bs, pending = [], [], None
        for mod in module:
            if isinstance(mod, nn.Linear):
                if pending is not None:
                    net.add_layers(Layer.dense(pending, 'linear'))
                ws.append(mod.weight.detach().cpu().double().numpy())
                bs.append(mod.bias.detach().cpu().double().numpy())
                pending = mod.out_features
            elif type(mod) in acts:
                net.add_layers(Layer.dense(pending, acts[type(mod)]))
                pending = None
            else:
                raise NotImplementedError(f"Module {type(mod).__name__} cannot be exported")
        net.setup()
        net.set_weights(ws, bs, input_scaling, output_scaling)
        return net

    def predict(self, X):
        """Numeric ``X`` of shape ``(n_features, m)`` -> ``(n_labels, m)``; symbolic ``X`` -> expressions."""
        if self._weights is None:
            raise RuntimeError("The network has no weights yet: call set_weights() or ANN.from_torch()")
        symbolic = sx.is_symbolic(X)
        h = sx.as_vector(X) if symbolic else np.atleast_2d(np.asarray(X, dtype=float))
        if symbolic and self.symbolic_mode == 'call' and len(self._layers) == 1 and \
                self._layers[0].activation.lower() in ('tanh', 'sigmoid', 'logistic') and \
                self._weights[0].shape[0] >= self.call_node_min_hidden:
            # one hidden layer: the network enters the expression graph as call nodes (a loop over the hidden units in
            # the generated device code, csrc/gp_device.cuh::mlp1_eval) with the z-score scalings folded into the weights
            W1, b1, W2, b2 = self._weights[0], self._bias[0], self._weights[1], self._bias[1]
            if self._in_mean is not None:
                b1 = b1 - W1 @ (self._in_mean / self._in_scale)
                W1 = W1 / self._in_scale[None, :]
            if self._out_mean is not None:
                W2 = W2 * self._out_scale[:, None]
                b2 = b2 * self._out_scale + self._out_mean
            mid = sx.register_mlp_table(W1, b1, W2, b2, self._layers[0].activation.lower())
            return sx.mlp_output_nodes(mid, h)
        if symbolic:
            if self._in_mean is not None:
                h = (h - self._in_mean) / self._in_scale
            for k, layer in enumerate(self._layers):
                y = self._weights[k] @ h + self._bias[k]
                h = np.array([_ACT[layer.activation.lower()](v) for v in y], dtype=object)
            y = self._weights[-1] @ h + self._bias[-1]
            if self._out_mean is not None:
                y = y * self._out_scale + self._out_mean
            return sx.as_vector(y)
        if self._in_mean is not None:
            h = (h - self._in_mean[:, None]) / self._in_scale[:, None]
        np_act = {'sigmoid': lambda v: 1. / (1. + np.exp(-v)), 'logistic': lambda v: 1. / (1. + np.exp(-v)),
                  'tanh': np.tanh, 'linear': lambda v: v, 'identity': lambda v: v,
                  'softplus': lambda v: np.log1p(np.exp(v)), 'soft_plus': lambda v: np.log1p(np.exp(v))}
        for k, layer in enumerate(self._layers):
            h = np_act[layer.activation.lower()](self._weights[k] @ h + self._bias[k][:, None])
        y = self._weights[-1] @ h + self._bias[-1][:, None]
        if self._out_mean is not None:
            y = y * self._out_scale[:, None] + self._out_mean[:, None]
        return y
