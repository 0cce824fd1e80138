"""Device timeline of the learner step graph (no nsys on the boxes): LearnerStep.enable_trace() inserts %globaltimer
marks at the phase boundaries, on whichever graph branch reaches them; this prints their mean offsets from the start of
the replay.  One GPU:
    python profiles/step_timeline.py [--ingest] [--no-prefetch] [--replays 200]
Data parallel (every rank prints its own table; clocks of different GPUs are not compared):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 \\
        profiles/step_timeline.py --prefetch
The marks cost one single-thread launch each (~2 us of extra serial work per mark on its branch): read the table for
the ORDER and the gaps between phases, not for the absolute step time (bench.py measures that without marks).
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from prism_b200.learner_step import LearnerStep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--replays", type=int, default=200)
    ap.add_argument("--capacity", type=int, default=1 << 18)
    ap.add_argument("--ingest", action="store_true", help="feed 4 new host steps per iteration (the e2e shape of the step)")
    ap.add_argument("--prefetch", dest="prefetch", action="store_true", default=None)
    ap.add_argument("--no-prefetch", dest="prefetch", action="store_false")
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    device = "cuda:%d" % local
    torch.cuda.set_device(local)
    pg = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=torch.device(device))
        pg = dist.group.WORLD
    cfg, agent, buf, trace, _ = bench.build_ours(rank, world, device, args.capacity, args.capacity, seed=4)
    if world > 1:
        dist.broadcast(agent.optimizer.arena, src=0)
        if agent.target_model is not None:
            dist.broadcast(agent.target_model._flat_arena, src=0)
    prefetch = (world == 1) if args.prefetch is None else args.prefetch
    step = LearnerStep(buf, agent, batch_size=bench.BATCH, use_cuda_graph=True, process_group=pg, rank=rank,
                       world_size=world, prefetch=prefetch)
    step.enable_trace()
    n_new = bench.STEPS_PER_ITER
    pre = trace.chunk((args.replays + 8) * n_new) if args.ingest else None
    rng = np.random.default_rng(2)
    u_host = torch.empty(step.B_global, dtype=torch.float64).pin_memory()

    def one(i):
        if not args.ingest:
            step.step()
            return
        sl = slice(i * n_new, (i + 1) * n_new)
        u_host.numpy()[:] = rng.random(step.B_global)
        torch.cuda.synchronize()                          # one pinned uniform buffer: never overwrite it in flight
        step.step(u=u_host, ingest=tuple(pre[k][sl] for k in ("stream", "obs", "action", "reward", "done", "trunc", "next_obs")))

    for i in range(8):
        one(i)
    acc = {}
    for i in range(args.replays):
        one(8 + i)
        for name, ns in step.trace_report().items():
            acc.setdefault(name, []).append(ns)
    lines = ["rank %d/%d  prefetch=%s ingest=%s exchange=%s  launches/step incl. %d marks: %s"
             % (rank, world, prefetch, args.ingest, getattr(step, "exchange", None), len(acc), step.launches_per_step)]
    for name, v in sorted(acc.items(), key=lambda kv: np.median(kv[1])):
        v = np.asarray(v, dtype=np.float64) / 1e3
        lines.append("  %-28s median %8.1f us   p10 %8.1f   p90 %8.1f" % (name, np.median(v), np.percentile(v, 10),
                                                                          np.percentile(v, 90)))
    if world > 1 and getattr(step, "peer", None) is not None:
        # phases INSIDE the one-launch gradient exchange (pb_peer_trace), a few replays, this rank's clock
        import ctypes as C
        from prism_b200 import _lib
        lib = _lib.load()
        out = (C.c_ulonglong * 16)()
        rows = []
        for i in range(10):
            _lib.check(lib.pb_peer_trace(1, out, 16), "pb_peer_trace")
            one(8 + args.replays + i)
            _lib.check(lib.pb_peer_trace(0, out, 16), "pb_peer_trace")
            if out[0] and out[4]:
                rows.append([(int(out[k]) - int(out[0])) / 1e3 for k in range(5)])
        if rows:
            med = np.median(np.asarray(rows), axis=0)
            lines.append("  inside the exchange kernel (us from CTA 0's start): handshake done %.1f | pulled + summed %.1f | "
                         "CTAs met %.1f | Adam applied %.1f" % (med[1], med[2], med[3], med[4]))
    if world > 1:
        for r in range(world):                            # one rank at a time
            if r == rank:
                print("\n".join(lines), flush=True)
            dist.barrier()
        dist.destroy_process_group()
    else:
        print("\n".join(lines))


if __name__ == "__main__":
    main()
