"""Generates tests/golden/reference_problems.json by running the UNMODIFIED
reference (imported in place from /root/reference through oracle/ref_import.py,
with its own C++ compiled by oracle/build_ref.sh) on the problems of
tests/problems.py.  Run in the build container only:

    python tests/golden/make_golden.py

For every problem it records what scipy SLSQP is given by
TrajectoryGenerator.generate_trajectory (x0, bounds, objective and SLSQP-ordered
constraint values at x0 and at a perturbed point, scipy's own 2-point
finite-difference Jacobian / gradient there) and, for the quick ones, the
OptimizeResult of the reference solve (x, status, nit, fun) together with the
reference's own reproducibility: the same solve started from x0 + k ulp, k = +/-1 .. +/-4
(`neighbours`), `status_stable` (all nine runs end with the same status) and
`stable` (status 0 every time and control points within 1e-5 of each other).
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


NEIGHBOUR_ULPS = (1, -1, 2, -2, 3, -3, 4, -4)     # starting points x0 + k ulp of the reproducibility runs


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
            constraints, cdl = getattr(gen, P + "__get_constraints")(N, wd, cc.derivative_constraints,
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
            # the reference's is_violation of a failed solve looks at the LAST constraint of its list only
            # (TG/trajectory_generator.py:252-261): recorded at both evaluation points
            pv = getattr(gen, P + "__print_violation")
            rec["last_block_violation_test"] = bool(pv(cdl[-1], xt))
            rec["last_block_violation_x0"] = bool(pv(cdl[-1], x0))
            rec["last_block_class"] = str(cdl[-1].constraint_class)
            if name in problems.SOLVE:
                captured = {}
                real_minimize = tgmod.minimize

                def spy(*a, **k):
                    captured["res"] = real_minimize(*a, **k)
                    # the reference against itself: the same call started one unit in the last place above / below x0
                    # (its forward differences amplify last-place noise by 1/h = 6.7e7, so long solves are not
                    # reproducible to 1e-5: the fixture records how far the reference lands from itself)
                    x0k = np.asarray(k["x0"], dtype=float)
                    captured["nbr"] = []
                    for ulps in NEIGHBOUR_ULPS:
                        xk = x0k.copy()
                        for _ in range(abs(ulps)):
                            xk = np.nextafter(xk, np.inf if ulps > 0 else -np.inf)
                        k2 = dict(k); k2["x0"] = xk
                        captured["nbr"].append(real_minimize(*a, **k2))
                    return captured["res"]
                tgmod.minimize = spy
                try:
                    cp, sf, viol = gen.generate_trajectory(cc, **kw)
                finally:
                    tgmod.minimize = real_minimize
                res = captured["res"]
                ncp = d * int(N) + 1           # control points and scale factor
                rec["solve"] = dict(x=np.asarray(res.x).tolist(), status=int(res.status), nit=int(res.nit),
                                    nfev=int(res.nfev), fun=float(res.fun), is_violation=bool(viol),
                                    scale_factor=float(sf), control_points=np.asarray(cp).tolist())
                nb = [dict(ulps=int(u), status=int(r.status), nit=int(r.nit), fun=float(r.fun),
                           dcp=float(np.abs(np.asarray(r.x)[:ncp] - np.asarray(res.x)[:ncp]).max()),
                           c_min_ineq=float(cons(r.x)[meq:].min()) if len(cons(r.x)) > meq else 0.0,
                           c_max_eq=float(np.abs(cons(r.x)[:meq]).max()) if meq else 0.0) for u, r in zip(NEIGHBOUR_ULPS, captured["nbr"])]
                rec["solve"]["neighbours"] = nb
                rec["solve"]["status_stable"] = bool(all(q["status"] == res.status for q in nb))
                rec["solve"]["stable"] = bool(rec["solve"]["status_stable"] and res.status == 0
                                              and all(q["dcp"] <= 1e-5 for q in nb))
                rec["solve"]["c_min_ineq"] = float(cons(res.x)[meq:].min()) if len(cons(res.x)) > meq else 0.0
                rec["solve"]["c_max_eq"] = float(np.abs(cons(res.x)[:meq]).max()) if meq else 0.0
        # JSON has no inf/nan: encode as strings
        out["problems"][name] = rec
        s = rec.get("solve")
        print("%-24s n=%2d meq=%2d m=%3d %s" % (name, rec["n"], rec["meq"], rec["m"],
                                               ("status=%d nit=%d stable=%s status_stable=%s nbr=%s" % (s["status"], s["nit"], s["stable"], s["status_stable"], (sorted(set(q["status"] for q in s["neighbours"])), min(q["nit"] for q in s["neighbours"]), max(q["nit"] for q in s["neighbours"]), "%.1e" % max(q["dcp"] for q in s["neighbours"]), "f %.9g..%.9g" % (min(q["fun"] for q in s["neighbours"]), max(q["fun"] for q in s["neighbours"]))))) if s else ""))

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
