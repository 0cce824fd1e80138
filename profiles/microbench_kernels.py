"""Warm, clock-ramped timing of individual libprism_b200 launches (CUDA events around a CUDA graph of
R back-to-back calls, so Python/ctypes launch cost is excluded).  Run on the GPU box."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prism_b200.agents import ops  # noqa: E402

dev = "cuda:0"


def graph_time(fn, reps=50, iters=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (reps * iters)


def main():
    torch.manual_seed(0)
    out = {}
    for (K, M, N, J) in [(1, 256, 256, 1024), (1, 256, 4, 256), (10, 64, 256, 1024), (1, 2048, 256, 1024),
                         (1, 32768, 512, 3136)]:
        x = torch.randn(M, J, device=dev)
        w = torch.randn(K, N, J, device=dev)
        b = torch.randn(K, N, device=dev)
        with torch.no_grad():
            t_mine = graph_time(lambda: ops.linear_heads(x, w, b, relu=True), reps=20 if M < 10000 else 3)
            t_torch = graph_time(lambda: torch.relu(torch.baddbmm(b.unsqueeze(1), x.unsqueeze(0).expand(K, M, J),
                                                                  w.transpose(1, 2))), reps=20 if M < 10000 else 3)
        flops = 2.0 * K * M * N * J
        out["linear_fwd K%d M%d N%d J%d" % (K, M, N, J)] = (t_mine, t_torch, flops / t_mine / 1e6)
    for k, (a, b_, tf) in out.items():
        print("%-40s mine %8.2f us   torch(cuBLAS+bias+relu) %8.2f us   mine %.2f TFLOP/s" % (k, a, b_, tf))
    # conv
    xo = (torch.rand(256, 10, 10, 6, device=dev) < 0.1).float()
    cw = torch.randn(16, 6, 3, 3, device=dev, requires_grad=True)
    cb = torch.randn(16, device=dev, requires_grad=True)
    with torch.no_grad():
        print("conv fwd B256: %.2f us" % graph_time(lambda: ops.conv3x3_relu_flatten(xo, cw, cb)))
    y = ops.conv3x3_relu_flatten(xo, cw, cb)
    gy = torch.randn_like(y)
    print("conv fwd+bwd B256: %.2f us" % graph_time(lambda: torch.autograd.grad(ops.conv3x3_relu_flatten(xo, cw, cb), (cw, cb), gy)))


if __name__ == "__main__" and "--tc" not in sys.argv:
    main()


def tc_bench():
    """tcgen05 3xTF32 GEMMs (forward, input gradient, weight gradient) vs the library SGEMM at the IQN / ensemble
    shapes of configs[0] and configs[4]."""
    shapes = [(1, 256, 256, 1024), (1, 512, 256, 1024), (1, 2048, 256, 1024), (1, 2048, 1024, 64), (1, 32768, 512, 3136), (1, 32768, 3136, 64), (10, 512, 512, 3136),
              (1, 8192, 1024, 1024)]
    for (K, M, N, J) in shapes:
        x = torch.randn(K, M, J, device=dev)
        w = torch.randn(K, N, J, device=dev)
        b = torch.randn(K, N, device=dev)
        dz = torch.randn(K, M, N, device=dev)
        y, dx, dw = torch.empty(K, M, N, device=dev), torch.empty(K, M, J, device=dev), torch.empty(K, N, J, device=dev)
        reps = 20 if M * N * J * K < 2e10 else 3
        fl = 2.0 * K * M * N * J
        rows = [
            ("fwd  ", lambda: ops.tc_gemm(y, x, 0, J, M * J, w, 0, J, N * J, K, M, N, J, bias=b, bias_bs=N, act=1),
             lambda: torch.relu(torch.baddbmm(b.unsqueeze(1), x, w.transpose(1, 2)))),
            ("dgrad", lambda: ops.tc_gemm(dx, dz, 0, N, M * N, w, 1, J, N * J, K, M, J, N), lambda: torch.bmm(dz, w)),
            ("wgrad", lambda: ops.tc_gemm(dw, dz, 1, N, M * N, x, 1, J, M * J, K, N, J, M),
             lambda: torch.bmm(dz.transpose(1, 2), x)),
        ]
        if J == 64 and K == 1:
            emb = torch.randn(512 if M >= 512 else 64, N, device=dev)
            q = M // emb.shape[0]
            rows.append(("phi*x", lambda: ops.tc_gemm(y, x, 0, J, M * J, w, 0, J, N * J, K, M, N, J, bias=b, bias_bs=N, act=1, mul=emb),
                         lambda: (torch.relu(torch.addmm(b[0], x[0], w[0].t())).view(q, emb.shape[0], N) * emb.unsqueeze(0))))
        for name, f_tc, f_lib in rows:
            with torch.no_grad():
                t_tc = graph_time(f_tc, reps=reps)
                t_lib = graph_time(f_lib, reps=reps)
            print("tc_gemm %s K%d M%d N%d J%d: tcgen05 3xTF32 %9.2f us (%6.1f TFLOP/s fp32-equivalent)   cuBLAS fp32 %9.2f us (%5.1f TFLOP/s)"
                  % (name, K, M, N, J, t_tc, fl / t_tc / 1e6, t_lib, fl / t_lib / 1e6), flush=True)


if __name__ == "__main__" and "--tc" in sys.argv:
    tc_bench()
