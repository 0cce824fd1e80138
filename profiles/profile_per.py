"""ncu driver for the PER kernels at BASELINE configs[2] size (2^24-leaf tree) and the Atari-shaped gather.
    python profiles/profile_per.py && ncu --set full --clock-control none --import-source on \
        -k regex:'tree_|upd_|store_gather' -c 20 -o gpurun_out/per_r01 python profiles/profile_per.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prism_b200 import PrioritizedTree, TransitionRing  # noqa: E402

dev = "cuda:0"
g = torch.Generator(device=dev)
g.manual_seed(1)
N = 1 << 24
tree = PrioritizedTree(N, device=dev, mode="stratified")
leaves = torch.empty(N, device=dev).exponential_(1.0, generator=g).add_(1e-8).sqrt_()
tree.build(leaves)                                                     # tree_reduce11 + tree_top
for B in (4096, 1 << 20):
    u = torch.rand(B, dtype=torch.float64, device=dev, generator=g)
    idx = torch.empty(B, dtype=torch.int64, device=dev)
    w = torch.empty(B, dtype=torch.float32, device=dev)
    prio = torch.rand(B, device=dev, generator=g)
    tree.sample(B, u=u, idx_out=idx, weight_out=w)                     # warp (4096) / thread (1M) descent
    tree.update_priority(idx, prio, sorted=True)                       # sparse (4096) / dense (1M) update
torch.cuda.synchronize()
del tree, leaves
cap, Bg = 1 << 17, 512
ring = TransitionRing(cap, (84, 84), frame_stack=4, n_step=3, gamma=0.99, storage_dtype=torch.uint8, obs_scale=True,
                      max_streams=8, staging_rows=64, device=dev)
ring.obs.random_(0, 256, generator=g)
seq = torch.arange(cap, device=dev)
ring.slot_seq.copy_(seq); ring.prev_link.copy_(seq - 1); ring.next_link.copy_(seq + 1)
ring.next_link[-1] = -1
ring.seq = cap
idx = torch.randint(8, cap - 8, (Bg,), device=dev, generator=g)
o = torch.empty(Bg, 4, 84, 84, device=dev); no = torch.empty_like(o)
r = torch.empty(Bg, 1, device=dev); gm = torch.empty(Bg, 1, device=dev)
nt = torch.empty(Bg, 1, dtype=torch.bool, device=dev); ac = torch.empty(Bg, 1, dtype=torch.int64, device=dev)
ring.gather(idx, o, no, r, gm, nt, ac)                                 # store_gather_kernel, Atari shapes
torch.cuda.synchronize()
print("profile_per ok")
