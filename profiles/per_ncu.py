"""Workload for `ncu --set full` of the priority-store kernels at BASELINE configs[2] (2^24 leaves):
one stratified batch of 4096 (sample + one-launch sorted write-back) and 64 batches in flight (one sampling launch +
scatter + streaming rebuild), plus the bulk build.  Each op runs twice warm before the profiled pass is reached by
ncu's --launch-skip (see profiles/README.md for the command line)."""
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prism_b200 import PrioritizedTree  # noqa: E402

dev = "cuda:0"
N, B = 1 << 24, 4096
g = torch.Generator(device=dev)
g.manual_seed(1)
tree = PrioritizedTree(N, device=dev, mode="stratified")
leaves = torch.empty(N, device=dev).exponential_(1.0, generator=g).add_(1e-8).sqrt_()
tree.build(leaves)
reps = int(os.environ.get("PER_NCU_REPS", "2"))
for K in (1, 16, 64):
    n = K * B
    u = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
    idx = torch.empty(n, dtype=torch.int64, device=dev)
    w = torch.empty(n, dtype=torch.float32, device=dev)
    prio = torch.rand(n, device=dev, generator=g)
    for _ in range(reps):
        tree.sample(B, u=u, idx_out=idx, weight_out=w, n_batches=K)
        tree.update_priority(idx, prio, sorted=(K == 1))
for _ in range(reps):
    tree.build(leaves)
torch.cuda.synchronize()
print("ok", tree.state_host())
