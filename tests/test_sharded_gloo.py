"""N > 1 host logic on CPU (world_size 2, gloo): shard stats all-gather -> virtual top -> stratum ownership.
Two ranks each hold an oracle shard; after a gloo all-gather of {p_sum, p_min, len} the union of what the
ranks sample must equal sampling ONE oracle tree over the concatenated leaves, and an averaged 'gradient'
all-reduce must equal the single-process mean (the data-parallel update rule of LearnerStep)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, cap, B, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.per_oracle import OracleTree
    from oracle.sharded_oracle import sample_rank
    rng = np.random.default_rng(9)
    leaves = np.sqrt(rng.exponential(1.0, world * cap).astype(np.float32))
    shard = OracleTree(cap)
    shard.build(leaves[rank * cap:(rank + 1) * cap])
    mine = torch.tensor([shard.sum[1], shard.min[1], float(len(shard))], dtype=torch.float32)
    gathered = [torch.zeros(3) for _ in range(world)]
    dist.all_gather(gathered, mine)
    all_psum = [float(g[0]) for g in gathered]
    all_pmin = [float(g[1]) for g in gathered]
    u = np.random.default_rng(10).random(B)            # every rank draws the same uniforms (same seed)
    ks, idx, w = sample_rank(shard, rank, all_psum, all_pmin, B, u)
    # data-parallel rule: local mean over the padded batch, scaled to the global mean, summed across ranks
    B_pad = B                                           # generous static bound for the test
    local = torch.zeros(4)
    local += torch.tensor(w[:len(ks)].sum()) * torch.ones(4) / B_pad       # stand-in "gradient": sum_b w_b / B_pad
    local *= B_pad / B                                                       # grad_scale = B_pad / B_glob
    dist.all_reduce(local, op=dist.ReduceOp.SUM)
    out_q.put((rank, ks, idx + rank * cap, w, local.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_sampling_equals_single_tree():
    world, cap, B = 2, 1 << 10, 256
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cap, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from oracle.per_oracle import OracleTree
    rng = np.random.default_rng(9)
    leaves = np.sqrt(rng.exponential(1.0, world * cap).astype(np.float32))
    big = OracleTree(world * cap)
    big.build(leaves)
    u = np.random.default_rng(10).random(B)
    oi, ow, _, _, _ = big.sample(u, 0.5, mode=1)
    got_idx = np.full(B, -1, np.int64)
    got_w = np.zeros(B, np.float32)
    runs = []
    for rank, ks, gidx, w, red in sorted(results, key=lambda r: r[0]):
        got_idx[ks] = gidx
        got_w[ks] = w
        runs.append(ks)
        assert np.array_equal(ks, np.arange(ks[0], ks[0] + len(ks)))      # every rank owns one contiguous run
        assert np.allclose(red, np.full(4, ow.sum() / B), rtol=1e-5)       # all-reduced value == global mean
    assert runs[0][-1] + 1 == runs[1][0] and runs[0][0] == 0 and runs[1][-1] == B - 1
    assert np.array_equal(got_idx, oi)
    assert np.allclose(got_w, ow, rtol=1e-6)
