"""Phase timeline of the priority-store kernels at BASELINE configs[2] (2^24 leaves, batches of 4096), warm, inside a
replayed CUDA graph: pb_tree_trace switches on %globaltimer marks inside the kernels (csrc/per_tree.cu); the table is
the LAST of three back-to-back sample -> write-back iterations of one replay, offsets from its sampling kernel's start.
    python profiles/per_phases.py [K ...]
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prism_b200 import PrioritizedTree, _lib  # noqa: E402

SLOTS = {0: "sample: CTA 0 starts", 1: "sample: top of the tree staged (last CTA)", 2: "sample: last CTA done",
         4: "chain: CTA 0 starts", 5: "chain: counters registered (last warp)", 6: "chain: leaves written",
         7: "chain: group nodes (L-5, L-10) written", 8: "chain: climbed (last warp)", 9: "chain: last-CTA ticket",
         10: "chain: top heap rebuilt", 11: "chain: state block written",
         16: "mark: CTA 0 starts", 17: "mark: last CTA done", 18: "leaf: CTA 0 starts", 19: "leaf: last CTA done",
         20: "sparse: CTA 0 starts", 21: "sparse: spans done (last CTA)", 22: "sparse: last-CTA ticket",
         23: "sparse: top heap rebuilt", 24: "sparse: state block written",
         28: "rebuild (last pass): CTA 0 starts", 29: "rebuild (last pass): lines done (last CTA)",
         30: "rebuild: last-CTA ticket", 31: "rebuild: top heap rebuilt", 32: "rebuild: state block written",
         36: "lines: CTA 0 starts", 37: "lines: last CTA done"}


def main():
    dev = "cuda:0"
    lib = _lib.load()
    Ks = [int(a) for a in sys.argv[1:]] or [1, 4, 16, 64, 256]
    N, B = 1 << 24, 4096
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    tree = PrioritizedTree(N, device=dev, mode="stratified")
    tree.build(torch.empty(N, device=dev).exponential_(1.0, generator=g).add_(1e-8).sqrt_())
    for K in Ks:
        n = K * B
        u = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
        idx = torch.empty(n, dtype=torch.int64, device=dev)
        w = torch.empty(n, dtype=torch.float32, device=dev)
        prio = torch.rand(n, device=dev, generator=g)

        def both():
            tree.sample(B, u=u, idx_out=idx, weight_out=w, n_batches=K)
            tree.update_priority(idx, prio, sorted=(K == 1))
        for _ in range(3):
            both()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(3):
                both()
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(20):
            graph.replay()
        ev1.record()
        torch.cuda.synchronize()
        per_iter = ev0.elapsed_time(ev1) / 60.0 * 1e3
        out = (C.c_ulonglong * 48)()
        _lib.check(lib.pb_tree_trace(1, out, 48), "pb_tree_trace")
        graph.replay()
        _lib.check(lib.pb_tree_trace(0, out, 48), "pb_tree_trace")
        marks = {k: int(out[k]) for k in SLOTS if out[k]}
        t0 = marks.get(0, min(marks.values()))
        print("K = %d batches of %d in flight: %.2f us per sample + write-back iteration (events, marks off)" % (K, B, per_iter))
        for k, t in sorted(marks.items(), key=lambda kv: kv[1]):
            print("   %8.2f us  %s" % ((t - t0) / 1e3, SLOTS[k]))
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
