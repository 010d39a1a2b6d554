"""Per-stage device time of one lock-step solve (CUDA events around whole solves; stage split from a launch list is
done with ncu + scripts/stage_times.py).  usage: stage_bench.py CONFIG BATCH [reps]"""
import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trajectory_generator_b200 import batch as tgb, synthetic as syn
name = sys.argv[1]; B = int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
bt = syn.make(name, B)
dev = torch.device("cuda:0")
par = torch.from_numpy(bt.par).to(dev); x0 = torch.from_numpy(bt.x0).to(dev)
bufs = tgb.SolveBuffers(bt.spec, B, dev)
ts = []
for r in range(reps + 1):
    x = x0.clone()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record(); out = tgb.solve(bt.spec, par, x, buffers=bufs); e.record(); torch.cuda.synchronize()
    if r: ts.append(s.elapsed_time(e))
print("%s %s B=%d: %.1f ms (min %.1f)  status0 %.4f nit %.2f  -> %.0f traj/s" % (os.environ.get("TG_LIB", "default").split("/")[-1], name, B, np.mean(ts), np.min(ts), (out["status"] == 0).float().mean().item(), out["nit"].float().mean().item(), B / np.min(ts) * 1e3))
