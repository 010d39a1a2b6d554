"""Generates tests/golden/sampling.json by running the UNMODIFIED reference samplers
(trajectory_generation/matrix_evaluation.py and spline_data_concatenater.py, imported in place through
oracle/ref_import.py) on a few cubic splines, among them the three of test_spline_data_concatenater.py:12-36.
Run in the build container only:   python tests/golden/make_golden_sampling.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_import  # noqa: E402


def main():
    ref_import.namespace()
    import trajectory_generation.matrix_evaluation as me
    from trajectory_generation.spline_data_concatenater import SplineDataConcatenater
    rng = np.random.default_rng(20261018)
    splines = {
        "concat1": (np.array([[4, 3, 7, 4, 8, 9, 5, 2, 7], [1, 2, 4, 6, 8, 9, 10, 13, 15]], dtype=float), 2.0),
        "concat2": (np.array([[5, 2, 7, 4, 8, 5], [10, 13, 15, 15, 11, 10]], dtype=float), 1.0),
        "concat3": (np.array([[4, 8, 5, 9, 2, 8, 13], [15, 11, 10, 8, 4, 2, 1]], dtype=float), 3.0),
        "rand3d": (rng.normal(size=(3, 11)) * 4, 0.7321),
        "rand2d_short": (rng.normal(size=(2, 4)), 1.9),
    }
    out = {"splines": {}}
    for name, (cps, scale) in splines.items():
        rec = {"control_points": cps.tolist(), "scale_factor": scale, "dataset": {}, "derivative_dataset": {}, "discrete": []}
        for num in (1, 2, 7, 100, 257):
            rec["dataset"][str(num)] = me.matrix_bspline_evaluation_for_dataset(3, cps, num).tolist()
        for r in (1, 2, 3):
            rec["derivative_dataset"][str(r)] = me.matrix_bspline_derivative_evaluation_for_dataset(3, r, scale, cps, 50).tolist()
        for (start, off, dt) in ((2.3, 0.0, 1.2), (0.0, 0.35, 0.1), (1.0, 0.0, 0.25)):
            for r in (0, 1, 2):
                if r == 0:
                    d, t, rem, end = me.matrix_bspline_evaluation_for_discrete_steps(3, cps, start, off, dt, scale)
                else:
                    d, t, rem, end = me.matrix_bspline_derivative_evaluation_for_discrete_steps(3, r, scale, cps, start, off, dt)
                rec["discrete"].append({"start_time": start, "offset": off, "dt": dt, "r": r, "data": d.tolist(),
                                        "time": t.tolist(), "remainder": rem, "end": end})
        tt = np.sort(rng.uniform(-0.5, scale * (cps.shape[1] - 3) + 0.5, size=40))
        rec["timedataset"] = {"time": tt.tolist(), "data": me.matrix_bspline_evaluation_for_timedataset(3, cps, tt, scale).tolist()}
        out["splines"][name] = rec
    # other orders (get_M_matrix serves 2 .. 5; order 1 raises in the reference) and the single-point helpers
    out["orders"] = {}
    for order in (2, 4, 5):
        cps = rng.normal(size=(2 + order % 2, order + 6)) * 3
        scale = 0.6 + 0.3 * order
        rec = {"control_points": cps.tolist(), "scale_factor": scale,
               "dataset": me.matrix_bspline_evaluation_for_dataset(order, cps, 41).tolist(), "derivative_dataset": {}, "discrete": []}
        for r in range(1, min(order, 3) + 1):
            rec["derivative_dataset"][str(r)] = me.matrix_bspline_derivative_evaluation_for_dataset(order, r, scale, cps, 37).tolist()
        for r in (0, 1):
            if r == 0:
                d, t, rem, end = me.matrix_bspline_evaluation_for_discrete_steps(order, cps, 0.5, 0.2, 0.3, scale)
            else:
                d, t, rem, end = me.matrix_bspline_derivative_evaluation_for_discrete_steps(order, r, scale, cps, 0.5, 0.2, 0.3)
            rec["discrete"].append({"start_time": 0.5, "offset": 0.2, "dt": 0.3, "r": r, "data": d.tolist(), "time": t.tolist(),
                                    "remainder": rem, "end": end})
        out["orders"][str(order)] = rec
    out["helpers"] = {"M": {str(o): me.get_M_matrix(o).tolist() for o in (2, 3, 4, 5)}, "points": []}
    for order in (2, 3, 4, 5):
        cp = rng.normal(size=(3, order + 1)) * 2
        for (t, tj, sf, r) in ((1.37, 1.0, 0.8, 0), (2.5, 2.0, 1.3, 1), (0.4, 0.0, 0.5, 2)):
            if r > order:
                continue
            val = (me.evaluate_point_on_interval(cp, t, tj, sf) if r == 0 else
                   me.evaluate_point_derivative_on_interval(cp, t, tj, sf, r))
            out["helpers"]["points"].append({"control_points": cp.tolist(), "t": t, "tj": tj, "scale": sf, "r": r,
                                             "value": np.asarray(val).tolist(),
                                             "T": (me.get_T_vector(order, t, tj, sf) if r == 0 else
                                                   me.get_T_derivative_vector(order, t, tj, r, sf)).tolist()})
    conc = SplineDataConcatenater(2)
    lst = [splines[k][0] for k in ("concat1", "concat2", "concat3")]
    sc = [splines[k][1] for k in ("concat1", "concat2", "concat3")]
    out["concatenate"] = []
    for r in (0, 1):
        d, t = conc.concatenate_spline_data(1.2, 2.3, [3, 3, 3], lst, sc, derivative_order=r)
        out["concatenate"].append({"dt": 1.2, "start_time": 2.3, "r": r, "data": d.tolist(), "time": t.tolist()})
    with open(os.path.join(HERE, "sampling.json"), "w") as f:
        json.dump(out, f)
    print("wrote sampling.json")


if __name__ == "__main__":
    main()
