"""Latency of small batches: lock-step stage kernels vs the fused persistent kernel.  usage: latency_bench.py CONFIG..."""
import sys, os, time, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trajectory_generator_b200 import batch as tgb, synthetic as syn
dev = torch.device("cuda:0")
for name in sys.argv[1:]:
    for B in (1, 16, 256, 1024, 4096, 16384):
        bt = syn.make(name, B)
        par = torch.from_numpy(bt.par).to(dev); x0 = torch.from_numpy(bt.x0).to(dev)
        bufs = tgb.SolveBuffers(bt.spec, B, dev)
        row = []
        for fused in (False, True):
            ts = []
            for r in range(4):
                x = x0.clone(); torch.cuda.synchronize()
                t = time.perf_counter()
                out = tgb.solve(bt.spec, par, x, jacobian="fd", buffers=bufs, fused=fused)
                torch.cuda.synchronize()
                if r: ts.append(time.perf_counter() - t)
            row.append((min(ts) * 1e3, out["x"].double().sum().item(), (out["status"] == 0).float().mean().item()))
        print("%-4s B=%6d  lock-step %8.2f ms   fused %8.2f ms   (status0 %.3f / %.3f)" % (name, B, row[0][0], row[1][0], row[0][2], row[1][2]), flush=True)
