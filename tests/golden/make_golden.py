"""Generates tests/golden/reference_problems.json by running the UNMODIFIED
reference (imported in place from /root/reference through oracle/ref_import.py,
with its own C++ compiled by oracle/build_ref.sh) on the problems of
tests/problems.py.  Run in the build container only:

    python tests/golden/make_golden.py

For every problem it records what scipy SLSQP is given by
TrajectoryGenerator.generate_trajectory (x0, bounds, objective and SLSQP-ordered
constraint values at x0 and at a perturbed point, scipy's own 2-point
finite-difference Jacobian / gradient there) and, for the quick ones, the
OptimizeResult of the reference solve (x, status, nit, fun).
"""
import contextlib
import io
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_import  # noqa: E402
import problems  # noqa: E402


def main():
    ns = ref_import.namespace()
    import scipy
    from scipy.optimize import Bounds
    from scipy.optimize._constraints import new_constraint_to_old
    from scipy.optimize._numdiff import approx_derivative
    import trajectory_generation.trajectory_generator as tgmod
    from trajectory_generation.objectives.objective_variables import (create_initial_objective_variables,
                                                                       create_objective_variable_bounds)
    eps = 1.4901161193847656e-08
    out = {"scipy_version": scipy.__version__, "numpy_version": np.__version__, "problems": {}}
    for name, make in problems.ALL.items():
        with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings(), np.errstate(all="ignore"):
            warnings.simplefilter("ignore")
            d, cc, kw = make(ns)
            gen = ns["TrajectoryGenerator"](d)
            P = "_TrajectoryGenerator"
            wd, sfc = cc.waypoint_constraints, cc.sfc_constraints
            mew0 = getattr(gen, P + "__get_num_intervals_free_space")(kw.get("num_intervals_free_space"), wd)
            nint = getattr(gen, P + "__get_num_intervals")(sfc, mew0, None)
            N = getattr(gen, P + "__get_num_control_points")(nint)
            seq = getattr(gen, P + "__get_point_sequence")(wd, sfc)
            constraints, _ = getattr(gen, P + "__get_constraints")(N, wd, cc.derivative_constraints,
                                                                   cc.turning_constraint, sfc, cc.obstacle_constraints)
            objective = getattr(gen, P + "__get_objective_function")(kw.get("objective_function_type",
                                                                             "minimal_velocity_and_time_path"))
            bnds = create_objective_variable_bounds(N, wd, d, 3)
            x0 = np.asarray(create_initial_objective_variables(N, seq, wd, d, 3, None, None), dtype=float)
            x0 = np.clip(x0, bnds.lb, bnds.ub)
            old = []
            for con in constraints:
                old += new_constraint_to_old(con, x0)
            eq = [c for c in old if c["type"] == "eq"]
            ineq = [c for c in old if c["type"] == "ineq"]

            def cons(x):
                parts = [np.atleast_1d(c["fun"](x)).ravel() for c in eq] + [np.atleast_1d(c["fun"](x)).ravel() for c in ineq]
                return np.concatenate(parts)

            def fobj(x):
                return np.atleast_1d(objective(x, N, d))

            meq = sum(len(np.atleast_1d(c["fun"](x0))) for c in eq)
            xt = problems.test_point(x0, d, N, seed=sum(map(ord, name)))
            xt = np.clip(xt, bnds.lb, bnds.ub)
            rec = dict(dimension=d, N=int(N), n=len(x0), meq=int(meq), m=int(len(cons(x0))),
                       x0=x0.tolist(), xl=np.asarray(bnds.lb).tolist(), xu=np.asarray(bnds.ub).tolist(),
                       f_x0=float(fobj(x0)[0]), c_x0=cons(x0).tolist(),
                       x_test=xt.tolist(), f_test=float(fobj(xt)[0]), c_test=cons(xt).tolist(),
                       grad_fd_test=np.atleast_1d(approx_derivative(fobj, xt, method="2-point", abs_step=eps,
                                                                    bounds=(bnds.lb, bnds.ub))).ravel().tolist(),
                       jac_fd_test=np.atleast_2d(approx_derivative(cons, xt, method="2-point", abs_step=eps,
                                                                   bounds=(bnds.lb, bnds.ub))).tolist())
            if name in problems.SOLVE:
                captured = {}
                real_minimize = tgmod.minimize

                def spy(*a, **k):
                    captured["res"] = real_minimize(*a, **k)
                    return captured["res"]
                tgmod.minimize = spy
                try:
                    cp, sf, viol = gen.generate_trajectory(cc, **kw)
                finally:
                    tgmod.minimize = real_minimize
                res = captured["res"]
                rec["solve"] = dict(x=np.asarray(res.x).tolist(), status=int(res.status), nit=int(res.nit),
                                    nfev=int(res.nfev), fun=float(res.fun), is_violation=bool(viol),
                                    scale_factor=float(sf), control_points=np.asarray(cp).tolist())
        # JSON has no inf/nan: encode as strings
        out["problems"][name] = rec
        s = rec.get("solve")
        print("%-24s n=%2d meq=%2d m=%3d %s" % (name, rec["n"], rec["meq"], rec["m"],
                                               ("status=%d nit=%d" % (s["status"], s["nit"])) if s else ""))

    def enc(o):
        if isinstance(o, float):
            if o != o: return "nan"
            if o in (float("inf"), float("-inf")): return "inf" if o > 0 else "-inf"
            return o
        if isinstance(o, list): return [enc(v) for v in o]
        if isinstance(o, dict): return {k: enc(v) for k, v in o.items()}
        return o
    with open(os.path.join(HERE, "reference_problems.json"), "w") as f:
        json.dump(enc(out), f)
    print("wrote", os.path.join(HERE, "reference_problems.json"))


if __name__ == "__main__":
    main()
