"""Minimal TensorDict stand-in for the static batch.

The reference builds a ``tensordict.TensorDict`` (prism/experience/timestep_buffer.py:84-97)
but only ever uses nested ``[]`` access, ``.clone()`` (prism/learner.py:71) and ``.device``
(prism/agents/agent.py:105).  tensordict is a third-party dependency we do not need.
"""
import torch


class Batch(dict):
    def __init__(self, data=None, batch_size=None, device=None):
        super().__init__()
        if data:
            for k, v in data.items():
                self[k] = Batch(v) if isinstance(v, dict) and not isinstance(v, Batch) else v
        self.batch_size = batch_size
        self._device = device

    @property
    def device(self):
        if self._device is not None:
            return torch.device(self._device)
        for v in self.values():
            if isinstance(v, (torch.Tensor, Batch)):
                return v.device
        return None

    def clone(self):
        out = Batch(batch_size=self.batch_size, device=self._device)
        for k, v in self.items():
            out[k] = v.clone() if isinstance(v, (torch.Tensor, Batch)) else v
        return out

    def to(self, device, non_blocking=False):
        out = Batch(batch_size=self.batch_size, device=device)
        for k, v in self.items():
            out[k] = v.to(device, non_blocking=non_blocking) if isinstance(v, (torch.Tensor, Batch)) else v
        return out
