"""Small solves / evaluations / sampling of every kernel family, for compute-sanitizer (memcheck, racecheck).
usage: compute-sanitizer --tool racecheck python scripts/sanitize.py"""
import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trajectory_generator_b200 import batch as tgb, synthetic as syn, matrix_evaluation as me
dev = torch.device("cuda:0")
for name, B in (("C2", 48), ("C3", 24), ("C4", 24), ("C5a", 16)):
    bt = syn.make(name, B); L = bt.layout
    par = torch.from_numpy(bt.par).to(dev)
    for mode in ("fd", "analytic"):
        x = torch.from_numpy(bt.x0).to(dev)
        out = tgb.solve(bt.spec, par, x, jacobian=mode, maxiter=12)
        torch.cuda.synchronize()
        print(name, mode, "status", np.unique(out["status"].cpu().numpy(), return_counts=True))
    xe = torch.from_numpy(syn.evaluation_points(bt)).to(dev)
    tgb.evaluate(bt.spec, par, xe)
    me.sample_batch((xe, L.d, L.N), num_points=70)
    me.sample_batch((xe, L.d, L.N), derivative_order=1, dt=0.3)
    torch.cuda.synchronize()
print("done")
