"""build_exp_buffer(config) -> TimestepBuffer, the seam of prism/factory/exp_buffer_factory.py:10-37:
the torchrl PrioritizedReplayBuffer(ListStorage) becomes a device-resident ring + sum/min trees."""
import torch

from ..experience import DevicePrioritizedReplayBuffer, TimestepBuffer


def build_exp_buffer(config):
    if getattr(config, "run_through_redis", False):
        # exp_buffer_factory.py:11-18: "server" = the learner / collector end, "client" = the buffer process
        if config.redis_side == "server":
            from ..async_components.async_experience_buffer import AsyncExperienceBufferInterface
            local = None
            if getattr(config, "redis_local_buffer", False):
                import copy
                local_config = copy.copy(config)
                local_config.run_through_redis = False
                local = build_exp_buffer(local_config)
            return AsyncExperienceBufferInterface(config.redis_host, config.redis_port, config.device, local_buffer=local)
        elif config.redis_side == "client":
            from ..async_components.async_experience_buffer import AsyncExperienceBuffer
            return AsyncExperienceBuffer(config.redis_host, config.redis_port)
    storage_dtype = torch.uint8 if getattr(config, "replay_storage_dtype", "float32") == "uint8" else torch.float32
    td_buffer = DevicePrioritizedReplayBuffer(
        capacity=config.experience_replay_capacity,
        alpha=config.per_alpha,
        beta=config.per_beta_start,
        batch_size=config.batch_size,
        device=config.device,
        prioritized=bool(config.use_per),
        sampling=getattr(config, "per_sampling", "iid"),
        storage_dtype=storage_dtype,
        obs_scale=getattr(config, "replay_obs_scale_255", False),
        max_streams=getattr(config, "replay_max_streams", 256),
        staging_rows=getattr(config, "replay_staging_rows", 256))
    return TimestepBuffer(td_buffer, frame_stack=config.frame_stack_size, device=config.device,
                          n_step=config.n_step_returns_length, gamma=config.gamma)
