"""build_exp_buffer(config): the seam of prism/factory/exp_buffer_factory.py:10-37.

Local mode: the torchrl ``PrioritizedReplayBuffer(ListStorage)`` / ``ReplayBuffer`` becomes a device-resident ring with
sum/min trees behind the same ``TimestepBuffer`` interface.  Redis mode (``run_through_redis``): the same two roles as
upstream -- ``redis_side == "server"`` is the learner / collector end, ``"client"`` the process that owns the buffer --
plus ``redis_local_buffer``: the learner keeps its own device buffer and only the collectors' steps cross the wire.
"""
import copy

import torch

from ..experience import DevicePrioritizedReplayBuffer, TimestepBuffer


def _device_buffer(cfg):
    opt = lambda name, default: getattr(cfg, name, default)            # fields the reference's Config does not have
    ring_and_trees = DevicePrioritizedReplayBuffer(
        cfg.experience_replay_capacity, alpha=cfg.per_alpha, beta=cfg.per_beta_start, batch_size=cfg.batch_size,
        device=cfg.device, prioritized=bool(cfg.use_per), sampling=opt("per_sampling", "iid"),
        storage_dtype=torch.uint8 if opt("replay_storage_dtype", "float32") == "uint8" else torch.float32,
        obs_scale=opt("replay_obs_scale_255", False), max_streams=opt("replay_max_streams", 256),
        staging_rows=opt("replay_staging_rows", 256))
    return TimestepBuffer(ring_and_trees, frame_stack=cfg.frame_stack_size, device=cfg.device,
                          n_step=cfg.n_step_returns_length, gamma=cfg.gamma)


def build_exp_buffer(config):
    if not getattr(config, "run_through_redis", False):
        return _device_buffer(config)
    from ..async_components import AsyncExperienceBuffer, AsyncExperienceBufferInterface
    side = config.redis_side
    if side == "client":
        return AsyncExperienceBuffer(config.redis_host, config.redis_port)
    if side != "server":
        raise ValueError("redis_side must be 'server' or 'client', got %r" % (side,))
    local = None
    if getattr(config, "redis_local_buffer", False):
        local_config = copy.copy(config)
        local_config.run_through_redis = False
        local = _device_buffer(local_config)
    return AsyncExperienceBufferInterface(config.redis_host, config.redis_port, config.device, local_buffer=local)
