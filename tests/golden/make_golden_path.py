"""Generates tests/golden/path_generator.json by running the UNMODIFIED reference's PathGenerator
(TG/path_generator.py, imported in place through oracle/ref_import.py) on tests/path_problems.py.
Run in the build container only:    python tests/golden/make_golden_path.py

Per problem: what scipy SLSQP is given (x0, bounds, objective and SLSQP-ordered constraint values at x0 and at a
perturbed point) and the OptimizeResult of the reference's own solve (x, status, nit, fun)."""
import contextlib
import importlib
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
import path_problems  # noqa: E402


NEIGHBOUR_ULPS = (1, -1, 2, -2, 3, -3, 4, -4)


def main():
    ns = ref_import.namespace()
    import scipy
    from scipy.optimize._constraints import new_constraint_to_old
    with contextlib.redirect_stdout(io.StringIO()):
        pgmod = importlib.import_module("trajectory_generation.path_generator")
    from trajectory_generation.objectives.objective_variables import (create_initial_objective_variables,
                                                                       create_objective_variable_bounds)
    out = {"scipy_version": scipy.__version__, "numpy_version": np.__version__, "problems": {}}
    for name, make in path_problems.ALL.items():
        with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings(), np.errstate(all="ignore"):
            warnings.simplefilter("ignore")
            d, cc, kw = make(ns)
            gen = pgmod.PathGenerator(d)
            P = "_PathGenerator"
            wd, sfc = cc.waypoint_constraints, cc.sfc_constraints
            mew0 = getattr(gen, P + "__get_num_intervals_free_space")(kw.get("num_intervals_free_space"), wd)
            nint = getattr(gen, P + "__get_num_intervals")(sfc, mew0, None)
            N = getattr(gen, P + "__get_num_control_points")(nint)
            seq = getattr(gen, P + "__get_point_sequence")(wd, sfc)
            constraints, _ = getattr(gen, P + "__get_constraints")(N, wd, cc.turning_constraint, sfc,
                                                                   cc.obstacle_constraints, kw.get("isIndirect", False))
            objective = getattr(gen, P + "__get_objective_function")(kw.get("objective_function_type", "minimal_velocity_path"))
            bnds = create_objective_variable_bounds(N, wd, d, 3)
            x0 = np.asarray(create_initial_objective_variables(N, seq, wd, d, 3, None, None), dtype=float)
            x0 = np.clip(x0, bnds.lb, bnds.ub)
            old = []
            for con in constraints:
                old += new_constraint_to_old(con, x0)
            eq = [c for c in old if c["type"] == "eq"]
            ineq = [c for c in old if c["type"] == "ineq"]

            def cons(x):
                return np.concatenate([np.atleast_1d(c["fun"](x)).ravel() for c in eq] +
                                      [np.atleast_1d(c["fun"](x)).ravel() for c in ineq])

            meq = sum(len(np.atleast_1d(c["fun"](x0))) for c in eq)
            xt = np.clip(problems.test_point(x0, d, N, seed=sum(map(ord, name))), bnds.lb, bnds.ub)
            rec = dict(dimension=d, N=int(N), n=len(x0), meq=int(meq), m=int(len(cons(x0))),
                       x0=x0.tolist(), xl=np.asarray(bnds.lb).tolist(), xu=np.asarray(bnds.ub).tolist(),
                       f_x0=float(objective(x0, N, d)), c_x0=cons(x0).tolist(),
                       x_test=xt.tolist(), f_test=float(objective(xt, N, d)), c_test=cons(xt).tolist())
            captured = {}
            real_minimize = pgmod.minimize

            def spy(*a, **k):
                captured["res"] = real_minimize(*a, **k)
                # the reference against itself from x0 + k ulp (see make_golden.py)
                x0k = np.asarray(k["x0"], dtype=float)
                captured["nbr"] = []
                for ulps in NEIGHBOUR_ULPS:
                    xk = x0k.copy()
                    for _ in range(abs(ulps)):
                        xk = np.nextafter(xk, np.inf if ulps > 0 else -np.inf)
                    k2 = dict(k); k2["x0"] = xk
                    captured["nbr"].append(real_minimize(*a, **k2))
                return captured["res"]
            pgmod.minimize = spy
            try:
                cp = gen.generate_path(cc, **kw)
            finally:
                pgmod.minimize = real_minimize
            res = captured["res"]
            rec["solve"] = dict(x=np.asarray(res.x).tolist(), status=int(res.status), nit=int(res.nit),
                                fun=float(res.fun), control_points=np.asarray(cp).tolist())
            ncp = d * int(N) + 1

            def feas(xv):
                c = cons(np.asarray(xv))
                return (float(c[meq:].min()) if len(c) > meq else 0.0, float(np.abs(c[:meq]).max()) if meq else 0.0)
            nb = [dict(ulps=int(u), status=int(r.status), nit=int(r.nit), fun=float(r.fun),
                       dcp=float(np.abs(np.asarray(r.x)[:ncp] - np.asarray(res.x)[:ncp]).max()),
                       c_min_ineq=feas(r.x)[0], c_max_eq=feas(r.x)[1]) for u, r in zip(NEIGHBOUR_ULPS, captured["nbr"])]
            rec["solve"]["neighbours"] = nb
            rec["solve"]["status_stable"] = bool(all(q["status"] == res.status for q in nb))
            rec["solve"]["stable"] = bool(rec["solve"]["status_stable"] and res.status == 0
                                          and all(q["dcp"] <= 1e-5 for q in nb))
            rec["solve"]["c_min_ineq"], rec["solve"]["c_max_eq"] = feas(res.x)
        out["problems"][name] = rec
        print("%-28s n=%2d meq=%2d m=%3d status=%d nit=%d stable=%s nbr dcp %.1e nit %d..%d" % (
            name, rec["n"], rec["meq"], rec["m"], rec["solve"]["status"], rec["solve"]["nit"], rec["solve"]["stable"],
            max(q["dcp"] for q in nb), min(q["nit"] for q in nb), max(q["nit"] for q in nb)))

    def enc(o):
        if isinstance(o, float):
            if o != o: return "nan"
            if o in (float("inf"), float("-inf")): return "inf" if o > 0 else "-inf"
            return o
        if isinstance(o, list): return [enc(v) for v in o]
        if isinstance(o, dict): return {k: enc(v) for k, v in o.items()}
        return o
    with open(os.path.join(HERE, "path_generator.json"), "w") as f:
        json.dump(enc(out), f)
    print("wrote", os.path.join(HERE, "path_generator.json"))


if __name__ == "__main__":
    main()
