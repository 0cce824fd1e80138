"""Host-side cost of the end-to-end loop (LearnerStep.step(u=host, ingest=host arrays)): host issue time per iteration
vs time including the GPU drain, plus a cProfile of the per-iteration path.  Run on a GPU box:
    python profiles/hostprof_e2e.py"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from prism_b200.learner_step import LearnerStep
dev = "cuda:0"
cfg, agent, buf, trace, _ = bench.build_ours(0, 1, dev, 1 << 18, 1 << 18, seed=4)
step = LearnerStep(buf, agent, batch_size=256, use_cuda_graph=True, prefetch=True)
for _ in range(5): step.step()
torch.cuda.synchronize()
N = 3000
pre = trace.chunk((N + 10) * 4)
u_host = torch.empty(256, dtype=torch.float64).pin_memory()
loss_host = torch.empty((), dtype=torch.float32).pin_memory()
def fused(i):
    sl = slice(i * 4, (i + 1) * 4)
    step.step(u=u_host, ingest=(pre["stream"][sl], pre["obs"][sl], pre["action"][sl], pre["reward"][sl], pre["done"][sl], pre["trunc"][sl], pre["next_obs"][sl]))
    step.copy_loss_to(loss_host)
def fused_nocopy(i):
    sl = slice(i * 4, (i + 1) * 4)
    step.step(ingest=(pre["stream"][sl], pre["obs"][sl], pre["action"][sl], pre["reward"][sl], pre["done"][sl], pre["trunc"][sl], pre["next_obs"][sl]))
import cProfile, pstats
for name, fn in [("fused step(u, ingest)+loss", fused), ("fused step(ingest) no u/loss", fused_nocopy), ("step() resident", lambda i: step.step())]:
    for i in range(5): fn(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(N): fn(i)
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print("%-28s host issue %.1f us/iter   incl. GPU drain %.1f us/iter" % (name, t_host / N * 1e6, t_all / N * 1e6))
pr = cProfile.Profile()
pr.enable()
for i in range(2000): fused(i)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
