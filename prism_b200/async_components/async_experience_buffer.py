"""The two ends of the Redis experience path, in front of the device buffer.

Reference (prism/async_components/async_experience_buffer.py:9-189, routed by exp_buffer_factory.py:11-18):

* ``AsyncExperienceBufferInterface`` -- what a collector and the learner hold instead of a buffer: ``extend`` ships
  completed steps to Redis in blocks of 100; ``sample`` takes a ready-made batch off Redis and copies it into the
  static batch.
* ``AsyncExperienceBuffer`` -- the process in the middle: pulls the step blocks, owns the real buffer, samples and ships
  batches back.  Upstream it samples uniformly and nobody ever calls ``update_priority``: the path bypasses PER.

Same classes, method names and wire bytes here.  What changes is what sits behind them:

* the middle process owns the DEVICE buffer (``build_exp_buffer``): blocks are decoded by the native codec into the
  arrays of ``TimestepBuffer.extend_batch`` (no per-step Python objects), batches leave through ``pack_numbers``;
* ``AsyncExperienceBufferInterface(..., local_buffer=buf)`` is the topology a device-resident buffer makes possible:
  the learner drains the step blocks into ITS OWN device buffer and samples there -- no batch ever crosses the wire,
  and prioritized sampling + ``update_priority`` work as on the local path.
"""
import time

import numpy as np
import torch

from .. import _lib
from . import wire
from .redis import RedisInterface


def _host(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


class AsyncExperienceBuffer(object):
    def __init__(self, redis_host, redis_port, redis_interface=None, experience_buffer=None):
        self._redis_interface = redis_interface or RedisInterface(redis_host, redis_port)
        self._batch_size = None
        self._experience_buffer = experience_buffer
        self._decoder = None
        self._time_between_command_pings = 1.0
        self._time_since_last_command_ping = 0.0
        self._n_collected = 0
        self._last_collect_call_timer = 0
        self._time_between_collect_calls = 0.1
        self._idle_sleep = 0.1

    def wait_for_config(self):
        config = self._redis_interface.get_config()
        while config is None:
            time.sleep(0.1)
            config = self._redis_interface.get_config()
        config.run_through_redis = False
        self._batch_size = config.batch_size
        if self._experience_buffer is None:
            from ..factory import exp_buffer_factory
            self._experience_buffer = exp_buffer_factory.build_exp_buffer(config)

    def _make_decoder(self):
        td = getattr(self._experience_buffer, "buffer", None)
        max_streams = getattr(td, "_storage_opts", {}).get("max_streams", 256) if td is not None else 256
        return wire.TimestepWireDecoder(max_streams=max_streams)

    def run(self, max_iterations=None):
        self.wait_for_config()
        running, it = True, 0
        while running and (max_iterations is None or it < max_iterations):
            it += 1
            self._get_latest_timesteps()
            if self._n_collected < self._batch_size:
                time.sleep(self._idle_sleep)
            else:
                self._experience_buffer.sample(batch_size=self._batch_size)
                self._transmit_batch()
            if time.perf_counter() - self._time_since_last_command_ping > self._time_between_command_pings:
                current_command = self._redis_interface.get_current_command()
                running = current_command != RedisInterface.SHUTDOWN_COMMAND
                self._time_since_last_command_ping = time.perf_counter()

    def _get_latest_timesteps(self):
        if time.perf_counter() - self._last_collect_call_timer < self._time_between_collect_calls:
            return
        if self._decoder is None:
            self._decoder = self._make_decoder()
        for flat in self._redis_interface.get_timestep_arrays():
            rows = self._decoder.feed(flat)
            if rows is not None:
                self._n_collected += self._experience_buffer.extend_batch(*rows)
        self._last_collect_call_timer = time.perf_counter()

    def _transmit_batch(self):
        b = self._experience_buffer
        tensors = [_host(t) for t in (b._obs, b._next_obs, b._reward, b._nonterminal, b._gamma, b._action)]
        self._redis_interface.add_batch_segments(wire.batch_segments(tensors))

    def _serialize_tensor(self, tensor):
        a = _host(tensor)
        return [a.ndim, *a.shape, int(a.size), *a.reshape(-1).tolist()]


class AsyncExperienceBufferInterface(object):
    def __init__(self, redis_host, redis_port, device, redis_interface=None, local_buffer=None, block_size=100):
        self._redis_interface = redis_interface or RedisInterface(redis_host, redis_port)
        self._batch = None
        self._batch_buffer = []
        self._timestep_buffer = []
        self._block_size = int(block_size)
        self.device = device
        self._obs = self._next_obs = self._reward = self._nonterminal = self._gamma = self._action = None
        self._local = local_buffer
        self._decoder = None
        self._poll_sleep = 0.01
        if local_buffer is not None:
            self.buffer = local_buffer.buffer            # learner.py:104-107 reaches buffer.buffer._sampler._beta

    @property
    def local_buffer(self):
        """The learner's own device buffer in local-buffer mode (hand THIS to ``LearnerStep`` and call ``drain()``
        between iterations), else None."""
        return self._local

    # ---- static batch ------------------------------------------------------------------------------------------------
    def set_static_batch(self, batch):
        if self._local is not None:
            self._local.set_static_batch(batch)
        self._batch = batch
        nxt = batch["next"]
        self._obs, self._next_obs, self._reward = batch["observation"], nxt["observation"], nxt["reward"]
        self._nonterminal, self._gamma, self._action = batch["nonterminal"], batch["gamma"], batch["action"]

    def get_static_batch(self):
        return self._local.get_static_batch() if self._local is not None else self._batch

    # ---- collector side ------------------------------------------------------------------------------------------------
    def extend(self, timestep):
        self._timestep_buffer.append(timestep)
        if len(self._timestep_buffer) >= self._block_size:
            self.flush()

    def flush(self):
        if self._timestep_buffer:
            self._redis_interface.submit_timesteps(self._timestep_buffer)
            self._timestep_buffer = []

    # ---- learner side ----------------------------------------------------------------------------------------------------
    def drain(self):
        """local_buffer mode: move every queued step block into the local device buffer; returns the steps added."""
        if self._decoder is None:
            td = self._local.buffer
            self._decoder = wire.TimestepWireDecoder(max_streams=getattr(td, "_storage_opts", {}).get("max_streams", 256))
        n = 0
        for flat in self._redis_interface.get_timestep_arrays():
            rows = self._decoder.feed(flat)
            if rows is not None:
                n += self._local.extend_batch(*rows)
        return n

    def sample(self, return_info=False, **kwargs):
        if self._local is not None:
            self.drain()
            return self._local.sample(return_info=return_info, **kwargs)
        arrays = self._redis_interface.get_waiting_batch_arrays()
        while arrays is None and len(self._batch_buffer) == 0:
            time.sleep(self._poll_sleep)
            arrays = self._redis_interface.get_waiting_batch_arrays()
        if arrays is not None:
            for flat in arrays:
                self._batch_buffer.append(self._deserialize_batch(flat))
        received = self._batch_buffer.pop(0)
        received = received[:5] + (received[5].long(),)
        for dst, src in zip((self._obs, self._next_obs, self._reward, self._nonterminal, self._gamma, self._action), received):
            dst.copy_(src, non_blocking=True)
        return (self._batch, 1) if return_info else self._batch

    def update_priority(self, indices, priorities):
        if self._local is None:
            raise _lib.PbError("priorities cannot be written back through the batch wire (the reference path bypasses "
                               "PER); construct the interface with local_buffer= to sample prioritized on the device")
        return self._local.update_priority(indices, priorities)

    def _deserialize_batch(self, flat):
        from ..experience.batch import Batch
        obs, next_obs, reward, nonterminal, gamma, action = (torch.from_numpy(a) for a in wire.split_batch(np.asarray(flat, dtype=np.float64)))
        if self._batch is None:
            dev = self.device
            B = obs.shape[0]
            batch = Batch({"observation": obs.to(dev),
                           "next": Batch({"observation": next_obs.to(dev), "reward": reward.to(dev)}, batch_size=B, device=dev),
                           "nonterminal": nonterminal.to(dev), "gamma": gamma.to(dev), "action": action.long().to(dev)},
                          batch_size=B, device=dev)
            self.set_static_batch(batch)
        return obs, next_obs, reward, nonterminal, gamma, action

    def empty(self):
        if self._local is not None:
            self._local.empty()

    def __len__(self):
        return len(self._local) if self._local is not None else 0
