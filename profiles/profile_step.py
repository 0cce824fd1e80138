"""Small driver for ncu: builds the bench's learner step on a reduced buffer (fewer set-up launches)
and replays it a few times.  Usage (on the GPU box):
    python profiles/profile_step.py && ncu --metrics gpu__time_duration.sum --clock-control none -c 800 \
        --csv --log-file gpurun_out/launches.csv python profiles/profile_step.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import prism_b200  # noqa: E402
from prism_b200.learner_step import LearnerStep  # noqa: E402

cap = int(os.environ.get("PB_PROFILE_CAPACITY", 1 << 20))
mode = os.environ.get("PB_PROFILE_MODE", "dqn")
dev = "cuda:0"
if mode == "dqn":
    cfg = prism_b200.minatar_dqn_per_config(device=dev, experience_replay_capacity=cap, batch_size=256,
                                            per_sampling="stratified", replay_max_streams=32,
                                            replay_staging_rows=65536, use_cuda_graph=False)
    obs_shape, A = bench.OBS_SHAPE, bench.N_ACTIONS
else:
    cfg = prism_b200.minatar_ids_iqn_config(device=dev, experience_replay_capacity=cap, batch_size=64,
                                            per_sampling="stratified", replay_max_streams=32,
                                            replay_staging_rows=65536, use_cuda_graph=False)
    obs_shape, A = (10, 10, 4), 3
if mode == "atari":
    # configs[4]: Atari-shaped 84x84x4 uint8 frames, IQN 64x64 + IDS, batch 512
    cap = min(cap, 1 << 16)
    cfg = prism_b200.atari_iqn_ids_config(device=dev, experience_replay_capacity=cap, per_sampling="stratified",
                                          replay_max_streams=32, replay_staging_rows=8192, replay_storage_dtype="uint8",
                                          replay_obs_scale_255=True, use_cuda_graph=False)
    torch.manual_seed(123)
    agent = prism_b200.build_agent(cfg, (cfg.frame_stack_size, 84, 84), 18)
    buf = prism_b200.build_exp_buffer(cfg)
    rng = np.random.default_rng(5)
    n = 8192
    frames = rng.integers(0, 256, (n + 32, 84 * 84), dtype=np.uint8)
    for rep in range(2):
        sid = (np.arange(n) % 32).astype(np.int32)
        buf.extend_batch(sid, frames[:n].reshape(n, 84, 84), rng.integers(0, 18, n).astype(np.int32),
                         (rng.random(n) < 0.05).astype(np.float32), rng.random(n) < (1 / 500), np.zeros(n, bool),
                         frames[32:32 + n].reshape(n, 84, 84))
    buf._flush()
    step = LearnerStep(buf, agent, batch_size=cfg.batch_size, use_cuda_graph=True, prefetch=True)
    for _ in range(4):
        step.step()
    torch.cuda.synchronize()
    print("profile_step (atari) ok: launches/step (ours)", step.launches_per_step)
    sys.exit(0)
torch.manual_seed(123)
agent = prism_b200.build_agent(cfg, obs_shape, A)
buf = prism_b200.build_exp_buffer(cfg)
trace = bench.Trace(4, int(np.prod(obs_shape)), A, 32)
n = min(cap, 65536)
c = trace.chunk(n)
buf.extend_batch(c["stream"], c["obs"].reshape((n,) + obs_shape), c["action"], c["reward"], c["done"], c["trunc"],
                 c["next_obs"].reshape((n,) + obs_shape))
buf._flush()
step = LearnerStep(buf, agent, batch_size=cfg.batch_size, use_cuda_graph=True, prefetch=True)
for _ in range(6):
    step.step()
torch.cuda.synchronize()
print("profile_step ok: launches/step (ours)", step.launches_per_step)
