"""ncu driver for the tcgen05 3xTF32 GEMM: the three GEMMs of the configs[4] IQN hidden layer
(M = 64 quantiles x 512 batch = 32768 rows, 3136 -> 512), forward / input gradient / weight gradient, and the
cos-embedding layer (64 -> 3136) with the fused phi (.) x epilogue.

    ncu --set full --clock-control none --import-source on -k regex:tc_gemm -c 4 -o gpurun_out/tc_r01 python profiles/profile_tc.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prism_b200.agents import ops  # noqa: E402

dev = "cuda:0"
M, N, J = 32768, 512, 3136
x = torch.randn(M, J, device=dev)
w = torch.randn(N, J, device=dev)
b = torch.randn(N, device=dev)
dz = torch.randn(M, N, device=dev)
y, dx, dw = torch.empty(M, N, device=dev), torch.empty(M, J, device=dev), torch.empty(N, J, device=dev)
basis = torch.randn(M, 64, device=dev)
wp, bp, emb = torch.randn(J, 64, device=dev), torch.randn(J, device=dev), torch.randn(512, J, device=dev)
for _ in range(2):
    ops.tc_gemm(y, x, 0, J, 0, w, 0, J, 0, 1, M, N, J, bias=b, act=1)                     # forward
    ops.tc_gemm(dx, dz, 0, N, 0, w, 1, J, 0, 1, M, J, N)                                    # input gradient
    ops.tc_gemm(dw, dz, 1, N, 0, x, 1, J, 0, 1, N, J, M)                                    # weight gradient
    ops.tc_gemm(dx, basis, 0, 64, 0, wp, 0, 64, 0, 1, M, J, 64, bias=bp, act=1, mul=emb)    # phi (.) x
torch.cuda.synchronize()
print("profile_tc ok")
