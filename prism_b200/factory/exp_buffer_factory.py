"""build_exp_buffer(config) -> TimestepBuffer, the seam of prism/factory/exp_buffer_factory.py:10-37:
the torchrl PrioritizedReplayBuffer(ListStorage) becomes a device-resident ring + sum/min trees."""
import torch

from ..experience import DevicePrioritizedReplayBuffer, TimestepBuffer


def build_exp_buffer(config):
    if getattr(config, "run_through_redis", False):
        raise NotImplementedError("the Redis transport bypasses PER and is outside the hot-path scope")
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
