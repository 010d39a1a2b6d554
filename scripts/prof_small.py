"""Launches the M1 evaluation kernel and the sampling kernel a few times on the C2 batch (for ncu captures)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trajectory_generator_b200 import batch as tgb, synthetic as syn, matrix_evaluation as me
name = sys.argv[1] if len(sys.argv) > 1 else "C2"; B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
bt = syn.make(name, B); L = bt.layout
dev = torch.device("cuda:0")
par = torch.from_numpy(bt.par).to(dev); x = torch.from_numpy(syn.evaluation_points(bt)).to(dev)
out = {}
samp = torch.empty((B, L.d, 512), dtype=torch.float64, device=dev)
for r in range(4):
    tgb.evaluate(bt.spec, par, x, out=out)
    me.sample_batch((x, L.d, L.N), num_points=512, out=samp)
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for q in range(10): me.sample_batch((x, L.d, L.N), num_points=512, out=samp)
    e.record(); torch.cuda.synchronize()
    print("10 sampling launches back to back: %.3f ms each" % (s.elapsed_time(e) / 10))
    s.record()
    for q in range(10): tgb.evaluate(bt.spec, par, x, out=out)
    e.record(); torch.cuda.synchronize()
    print("10 eval launches back to back: %.3f ms each" % (s.elapsed_time(e) / 10))
