"""Where do the FD-emulation solve (host build of the kernel source) and the reference's scipy solve separate?
usage: fd_parity_probe.py CONFIG COUNT"""
import sys, os, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import hostsim_loader, tg_oracle
from trajectory_generator_b200 import synthetic as syn
from trajectory_generator_b200.problem import pack_problem
from scipy.optimize import minimize, Bounds
name = sys.argv[1]; count = int(sys.argv[2])
hs = hostsim_loader.load()
b = syn.make(name, max(64, count))
for i in range(count):
    d, cc, kw = syn.container_for(b, i)
    obj = kw.get("objective_function_type", syn.OBJECTIVE[name])
    pp = pack_problem(d, cc, obj, kw.get("num_intervals_free_space"))
    op = tg_oracle.OracleProblem(d, cc, obj, kw.get("num_intervals_free_space"))
    calls = []
    def fun(x):
        calls.append(np.array(x)); return op.fun(x)
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore")
        res = minimize(fun, x0=op.x0, method="SLSQP", bounds=Bounds(lb=op.xl, ub=op.xu), constraints=op.scipy_constraints(),
                       options={"disp": False, "maxiter": 100, "ftol": 1e-6})
    # accepted iterates = points at which scipy starts a forward-difference sweep (next call differs in one entry by ~1.5e-8)
    its = []
    for a, bb in zip(calls[:-1], calls[1:]):
        dd = np.abs(a - bb)
        if (dd > 0).sum() == 1 and dd.max() < 1e-7 and (not its or np.abs(its[-1] - a).max() > 0):
            its.append(a)
    its = its[1:]          # the first sweep is at x0
    # the reference against itself: same solve from x0 moved by one unit in the last place
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore")
        res2 = minimize(op.fun, x0=np.nextafter(op.x0, np.inf), method="SLSQP", bounds=Bounds(lb=op.xl, ub=op.xu), constraints=op.scipy_constraints(),
                        options={"disp": False, "maxiter": 100, "ftol": 1e-6})
    mine = hs.solve(pp, fd=True, trace=True)
    k = pp.layout.ia + 1
    first = None
    for it, xk in enumerate(its):
        if it < mine["nit"]:
            dd = np.abs(mine["trace"][it][2:] - xk).max()
            if dd > 1e-7 and first is None: first = (it + 1, float(dd))
    print("%3d ref st %d nit %3d | mine st %d nit %3d | final dx %.1e | first >1e-7 at %s | ref vs ref(x0+1ulp): st %d nit %3d dx %.1e" % (i, res.status, res.nit, mine["status"], mine["nit"], np.abs(mine["x"][:k] - res.x[:k]).max(), first, res2.status, res2.nit, np.abs(res2.x[:k] - res.x[:k]).max()))
