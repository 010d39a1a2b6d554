"""Device time of one lock-step solve and its split into the two stage kernels (CUDA events recorded by the library).
usage: stage_bench.py CONFIG BATCH [reps] [fd|analytic]      (TG_LIB=<variant .so> selects a tuning build)"""
import sys, os, ctypes, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trajectory_generator_b200 import batch as tgb, synthetic as syn, _native
name = sys.argv[1]; B = int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mode = sys.argv[4] if len(sys.argv) > 4 else "fd"
bt = syn.make(name, B)
dev = torch.device("cuda:0")
par = torch.from_numpy(bt.par).to(dev); x0 = torch.from_numpy(bt.x0).to(dev)
bufs = tgb.SolveBuffers(bt.spec, B, dev)
ts = []
for r in range(reps + 1):
    x = x0.clone()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record(); out = tgb.solve(bt.spec, par, x, jacobian=mode, buffers=bufs); e.record(); torch.cuda.synchronize()
    if r: ts.append(s.elapsed_time(e))
lib = _native.lib()
lib.tg_set_stage_timing(1)
x = x0.clone(); tgb.solve(bt.spec, par, x, jacobian=mode, buffers=bufs); torch.cuda.synchronize()
lib.tg_set_stage_timing(0)
st = (ctypes.c_double * 6)(); lib.tg_last_solve_stats(st, 6)
print("%-8s %s %s B=%d: %.1f ms (min %.1f) ls %.1f qp %.1f | status0 %.4f nit %.2f -> %.0f traj/s" % (
    os.environ.get("TG_LIB", "default").split("/")[-1], name, mode, B, np.mean(ts), np.min(ts), st[0], st[1],
    (out["status"] == 0).float().mean().item(), out["nit"].float().mean().item(), B / np.min(ts) * 1e3))
