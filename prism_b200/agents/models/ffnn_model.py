"""FFNNModel -- MLP with optional pre-layer LayerNorm.

Constructor arguments and the ``model.<i>`` parameter names match the reference class
(prism/agents/models/ffnn_model.py:46-97) so that state_dicts are interchangeable; the
layer stack is: [LN(in)] Linear act ... [LN] Linear [output_act], evaluated by ops.run_sequential
(fused Linear+ReLU launches, tcgen05 GEMMs and fused LayerNorm kernels for large activations).
"""
import numpy as np
import torch
import torch.nn as nn


def _mlp_stack(widths, use_layer_norm, norm_first, act_fn, output_act_fn):
    """Yield the modules of an MLP whose Linear layers map widths[i] -> widths[i+1]."""
    last = len(widths) - 2
    for i, (fan_in, fan_out) in enumerate(zip(widths[:-1], widths[1:])):
        if use_layer_norm and (i > 0 or norm_first):
            yield nn.LayerNorm(fan_in)
        yield nn.Linear(fan_in, fan_out)
        if i != last:
            yield act_fn()
    if output_act_fn is not None:
        yield output_act_fn()


class FFNNModel(nn.Module):
    def __init__(self, n_input_features, n_output_features, n_layers, layer_width, use_layer_norm,
                 use_p_norm=False, apply_layer_norm_first_layer=True, output_act_fn=None, act_fn=nn.ReLU,
                 sparse_init_p=0.0, device="cpu"):
        super().__init__()
        if use_p_norm:
            raise NotImplementedError("PNorm is not used by any reference configuration")
        self.device = device
        widths = [n_input_features] + [layer_width] * (n_layers - 1) + [n_output_features]
        self.model = nn.Sequential(*_mlp_stack(widths, use_layer_norm, apply_layer_norm_first_layer,
                                               act_fn, output_act_fn)).to(device)
        if sparse_init_p > 0.0:
            self._sparse_init(sparse_init_p)

    @torch.no_grad()
    def _sparse_init(self, p_zero):
        # xavier weights with a Bernoulli(1-p) keep-mask and zero biases (ffnn_model.py:82-88)
        for lin in (m for m in self.model if isinstance(m, nn.Linear)):
            nn.init.xavier_uniform_(lin.weight)
            nn.init.zeros_(lin.bias)
            lin.weight.mul_(torch.bernoulli(torch.full_like(lin.weight, 1.0 - p_zero)))

    def forward(self, x):
        if not isinstance(x, torch.Tensor):
            x = torch.from_numpy(np.asarray(x, dtype=np.float32)).to(self.device)
        x = x.reshape(x.shape[0], -1)
        if x.is_cuda:
            from .. import ops
            return ops.run_sequential(self.model, x)     # Linear(+ReLU) layers as fused launches
        return self.model(x)

    def log(self, logger):
        pass
