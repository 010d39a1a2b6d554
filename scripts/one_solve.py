import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from trajectory_generator_b200 import batch as tgb, synthetic as syn
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else syn.FULL_BATCH[name]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
bt = syn.make(name, B)
dev = torch.device("cuda:0")
par = torch.from_numpy(bt.par).to(dev); x0 = torch.from_numpy(bt.x0).to(dev)
bufs = tgb.SolveBuffers(bt.spec, B, dev)
for r in range(reps):
    x = x0.clone()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record(); out = tgb.solve(bt.spec, par, x, buffers=bufs, jacobian=(sys.argv[4] if len(sys.argv) > 4 else "fd")); e.record(); torch.cuda.synchronize()
    print("solve %d: %.1f ms, status0 %.3f mean nit %.1f" % (r, s.elapsed_time(e), (out["status"] == 0).float().mean().item(), out["nit"].float().mean().item()))
xe = torch.from_numpy(syn.evaluation_points(bt)).to(dev)
o = tgb.evaluate(bt.spec, par, xe); torch.cuda.synchronize()
for r in range(reps + 1):
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record(); tgb.evaluate(bt.spec, par, xe, out=o); e.record(); torch.cuda.synchronize()
    print("eval %d: %.3f ms  (%.1f M evaluations/s)" % (r, s.elapsed_time(e), B / s.elapsed_time(e) / 1e3))
