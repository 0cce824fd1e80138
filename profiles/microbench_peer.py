"""Per-operation cost of the data-parallel exchange steps, peer-memory kernels vs NCCL (run under torchrun):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 profiles/microbench_peer.py
Every op is captured in a CUDA graph of REPS back-to-back calls and timed with CUDA events (max over ranks)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prism_b200 import _lib  # noqa: E402
from prism_b200.peer import PeerGroup  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda:%d" % local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
REPS = 200


class Opt:
    def __init__(self, n):
        self.arena = torch.randn(n, device=dev)
        self.exp_avg = torch.zeros(n, device=dev)
        self.exp_avg_sq = torch.zeros(n, device=dev)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)
        self.partials = torch.zeros(4096, device=dev)
        self.norm_out = torch.zeros(2, device=dev)
        self.lr, self.betas, self.eps, self.max_grad_norm = 1e-4, (0.9, 0.999), 1.5e-4, 10.0


def timed(name, fn, note=""):
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    dist.barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REPS):
            fn()
    g.replay()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) * 1e3 / (5 * REPS)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("%-46s %8.2f us  %s" % (name, float(t), note), flush=True)
    del g


for n in (2_378_000, 18_167_208):
    n = (n + 3) // 4 * 4
    peer = PeerGroup.create(dist.group.WORLD, rank, world, n, dev)
    opt = Opt(n)
    peer.grad.normal_()
    state = torch.zeros(64, dtype=torch.uint8, device=dev)
    gathered = torch.zeros(world, 64, dtype=torch.uint8, device=dev)
    flat = torch.randn(n, device=dev)
    lib = _lib.load()
    st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
    if rank == 0:
        print("== world %d, n = %d parameters (%.1f MB)" % (world, n, n * 4 / 1e6), flush=True)
    timed("peer barrier", peer.barrier)
    timed("peer state all-gather (64 B/rank)", lambda: peer.state_allgather(state))
    timed("nccl all_gather_into_tensor (64 B/rank)", lambda: dist.all_gather_into_tensor(gathered.view(-1), state))
    timed("peer all-reduce + clip + Adam (fused)", lambda: peer.allreduce_adam(opt))
    def nccl_path():
        dist.all_reduce(flat)
        _lib.check(lib.pb_adam_clip_step(n, opt.arena.data_ptr(), flat.data_ptr(), opt.exp_avg.data_ptr(),
                                         opt.exp_avg_sq.data_ptr(), opt.step_count.data_ptr(), opt.lr, 0.9, 0.999, opt.eps,
                                         opt.max_grad_norm, opt.norm_out.data_ptr(), opt.partials.data_ptr(), st()), "adam")
    timed("nccl all_reduce + sumsq + clip + Adam", nccl_path)
    timed("local sumsq + clip + Adam only (no exchange)",
          lambda: _lib.check(lib.pb_adam_clip_step(n, opt.arena.data_ptr(), flat.data_ptr(), opt.exp_avg.data_ptr(),
                                                   opt.exp_avg_sq.data_ptr(), opt.step_count.data_ptr(), opt.lr, 0.9, 0.999,
                                                   opt.eps, opt.max_grad_norm, opt.norm_out.data_ptr(),
                                                   opt.partials.data_ptr(), st()), "adam"))
    del peer, opt
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
