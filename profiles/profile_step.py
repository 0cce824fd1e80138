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
torch.manual_seed(123)
agent = prism_b200.build_agent(cfg, obs_shape, A)
buf = prism_b200.build_exp_buffer(cfg)
trace = bench.Trace(4, int(np.prod(obs_shape)), A, 32)
n = min(cap, 65536)
c = trace.chunk(n)
buf.extend_batch(c["stream"], c["obs"].reshape((n,) + obs_shape), c["action"], c["reward"], c["done"], c["trunc"],
                 c["next_obs"].reshape((n,) + obs_shape))
buf._flush()
step = LearnerStep(buf, agent, batch_size=cfg.batch_size, use_cuda_graph=True)
for _ in range(6):
    step.step()
torch.cuda.synchronize()
print("profile_step ok: launches/step (ours)", step.launches_per_step)
