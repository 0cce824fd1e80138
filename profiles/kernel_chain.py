"""Warm, in-graph cost of the configs[1] step's kernels one by one (no nsys on the boxes; ncu's per-launch times are
cold-cache): each op is captured N times back to back in one CUDA graph on one stream and replayed; the figure is the
replay time / N, i.e. kernel duration + the launch gap of a dependent chain.  Library GEMM beside ours for scale.
    python profiles/kernel_chain.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prism_b200.agents import ops  # noqa: E402

DEV = "cuda:0"
N = 40


def timed(name, fn, reps=30):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(N):
            fn()
    for _ in range(5):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print("%-58s %7.2f us per call" % (name, e0.elapsed_time(e1) * 1e3 / (reps * N)))


def main():
    torch.manual_seed(0)
    B, C = 256, 6
    obs = torch.rand(B, 10, 10, C, device=DEV)
    cw, cb = torch.randn(16, C, 3, 3, device=DEV) * 0.1, torch.zeros(16, device=DEV)
    w1, b1 = torch.randn(1, 256, 1024, device=DEV) * 0.03, torch.zeros(1, 256, device=DEV)
    w2, b2 = torch.randn(1, 6, 256, device=DEV) * 0.05, torch.zeros(1, 6, device=DEV)
    with torch.no_grad():
        emb = ops.conv3x3_relu_flatten(obs, cw, cb)
        h = ops.linear_heads(emb, w1, b1, relu=True)
        timed("conv3x3+relu+flatten fwd (256 x 10x10x6 -> 1024)", lambda: ops.conv3x3_relu_flatten(obs, cw, cb))
        timed("linear 1024 -> 256 + relu fwd (ours)", lambda: ops.linear_heads(emb, w1, b1, relu=True))
        timed("linear 1024 -> 256 + relu fwd (library addmm + relu)",
              lambda: torch.relu(torch.addmm(b1[0], emb, w1[0].t())))
        timed("narrow 256 -> 6 fwd", lambda: ops.linear_heads(h[0] if h.dim() == 3 else h, w2, b2))
    # phases inside ONE warm launch of the forward GEMM (pb_gemm_trace)
    import ctypes as C
    from prism_b200 import _lib
    lib = _lib.load()
    names = ["first chunk staged", "K loop done", "partial tile written", "cluster barrier", "split-K sum + epilogue",
             "exit barrier"]
    for label, fn in (("forward 256 x 256 x 1024", lambda: ops.linear_heads(emb, w1, b1, relu=True)),):
        with torch.no_grad():
            for _ in range(3):
                fn()
            lib.pb_gemm_trace(1, None, 0)
            fn()
        out = (C.c_ulonglong * 8)()
        lib.pb_gemm_trace(0, out, 8)
        print("phases of one launch, %s (latest CTA, us after the earliest CTA start):" % label)
        for k, nm in enumerate(names):
            print("   %6.2f  %s" % ((out[k + 1] - out[0]) / 1e3, nm))
    # backward pieces
    embg = emb.clone().requires_grad_(False)
    w1g, b1g = w1.clone().requires_grad_(True), b1.clone().requires_grad_(True)
    embr = emb.clone().requires_grad_(True)
    gy = torch.randn(1, 256, 256, device=DEV)

    def lin_fb():
        y = ops.linear_heads(embr, w1g, b1g, relu=True)
        torch.autograd.grad(y, (embr, w1g, b1g), gy)
    timed("linear fwd + dgrad + wgrad (two branches)", lin_fb)
    cwg, cbg = cw.clone().requires_grad_(True), cb.clone().requires_grad_(True)
    ge = torch.randn(B, 1024, device=DEV)

    def conv_fb():
        y = ops.conv3x3_relu_flatten(obs, cwg, cbg)
        torch.autograd.grad(y, (cwg, cbg), ge)
    timed("conv fwd + bwd (mma) + reduce", conv_fb)


if __name__ == "__main__":
    main()
