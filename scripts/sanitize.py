"""Small solves / evaluations / sampling of every kernel family, for compute-sanitizer (memcheck, racecheck).
usage: compute-sanitizer --tool memcheck python scripts/sanitize.py     (ONE tool per gpurun call)
Covers: the fixed-shape and the generic stage kernels (TG_GENERIC_KERNELS), both Jacobian modes, the fused kernel, M1,
the samplers of every order, the point helper, the builders (initial guess, boxes, intervals per corridor), the
mixed-shape device solve, the spline order converter and the 24 legacy symbols."""
import sys, os, json, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from trajectory_generator_b200 import batch as tgb, synthetic as syn, matrix_evaluation as me, _native
from trajectory_generator_b200.batched import CorridorProblems
from trajectory_generator_b200.spline_order_converter import SmoothingSpline
dev = torch.device("cuda:0")
for generic in (False, True):
    if generic:
        os.environ["TG_GENERIC_KERNELS"] = "1"
    for name, B in (("C2", 48), ("C3", 24), ("C4", 24), ("C5a", 16), ("C5c", 16)):
        bt = syn.make(name, B); L = bt.layout
        par = torch.from_numpy(bt.par).to(dev)
        for mode in ("fd", "analytic"):
            x = torch.from_numpy(bt.x0).to(dev)
            out = tgb.solve(bt.spec, par, x, jacobian=mode, maxiter=12)
            torch.cuda.synchronize()
            print(name, "generic" if generic else "fixed", mode, "status", np.unique(out["status"].cpu().numpy(), return_counts=True))
        xe = torch.from_numpy(syn.evaluation_points(bt)).to(dev)
        tgb.evaluate(bt.spec, par, xe)
        tgb.evaluate(bt.spec, par, xe, want=("f", "jnl"))
        torch.cuda.synchronize()
os.environ.pop("TG_GENERIC_KERNELS", None)
bt = syn.make("C2", 8)
tgb.solve(bt.spec, torch.from_numpy(bt.par).to(dev), torch.from_numpy(bt.x0).to(dev), maxiter=5, fused=True)
rng = np.random.default_rng(0)
for order in (2, 3, 4, 5):
    cps = torch.from_numpy(rng.normal(size=(5, 3, order + 6))).to(dev)
    sf = torch.full((5,), 0.8, dtype=torch.float64, device=dev)
    me.sample_batch(cps, num_points=71, order=order)
    me.sample_batch(cps, sf, derivative_order=1, dt=0.31, order=order)
    pc = torch.from_numpy(rng.normal(size=(7, 2, order + 1))).to(dev)
    t = torch.full((7,), 1.3, dtype=torch.float64, device=dev)
    me.interval_points_batch(pc, t, t - 0.3, t * 0 + 0.9, 1)
pts = np.cumsum(rng.normal(size=(40, 3, 4)) * 4, 2)
pads = rng.uniform(2, 3, (40, 3, 3))
v0 = pts[:, :, 1] - pts[:, :, 0]; v0 /= np.linalg.norm(v0, 2, 1)[:, None]
tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
cp = CorridorProblems(3, corridor_points=tt(pts), corridor_pads=tt(pads), start_velocity=tt(v0), end_zero_velocity=True,
                      max_velocity=5.0, max_acceleration=0.3, objective_function_type="minimal_velocity_path")
out = cp.solve(maxiter=6)
print("corridor shapes", cp.shapes())
G = json.load(open(os.path.join(ROOT, "tests", "golden", "smoothing.json")))["cases"]["cubic_to_cubic"]
sm = SmoothingSpline(G["new_order"], 2, G["resolution"])
sm.generate_new_control_points(np.array(G["cp"], dtype=float), G["scale"], G["old_order"])
print("smoothing", sm.last_result)
import legacy_abi
for k in json.load(open(os.path.join(ROOT, "tests", "golden", "native_kats.json")))["kats"]:
    legacy_abi.call(_native.lib(), k)
torch.cuda.synchronize()
print("done")
