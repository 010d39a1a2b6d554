"""Generates tests/golden/build.json by running the UNMODIFIED reference's problem-construction helpers
(create_initial_objective_variables, TG/objectives/objective_variables.py; get2D/3DRotationAndTranslationFromPoints
and SFC.getRotatedBounds, DS/safe_flight_corridor.py) through oracle/ref_import.py.
Run in the build container only:   python tests/golden/make_golden_build.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_import  # noqa: E402


def main():
    ns = ref_import.namespace()
    from trajectory_generation.objectives.objective_variables import create_initial_objective_variables
    import trajectory_generation.constraint_data_structures.safe_flight_corridor as sfcmod
    W, WD = ns["Waypoint"], ns["WaypointData"]
    rng = np.random.default_rng(20261018)
    col = lambda v: np.asarray(v, dtype=float).reshape(-1, 1)
    out = {"initial": [], "boxes": []}
    # initial variables: straight lines and polylines (incl. a commensurate one whose control points fall on vertices)
    cases = []
    for d in (2, 3):
        for N in (8, 11, 17):
            cases.append((d, N, rng.normal(size=(d, 2)) * 5, None))
            cases.append((d, N, np.cumsum(rng.normal(size=(d, 5)) * 3, 1), None))
    cases.append((2, 9, np.array([[0.0, 2.0, 4.0, 6.0, 8.0], [0.0, 0.0, 0.0, 0.0, 0.0]]), None))
    cases.append((2, 13, np.array([[0.0, 3.0, 3.0, 0.0], [0.0, 0.0, 4.0, 4.0]]), None))
    # intermediate waypoints: 1 and 2 of them, with velocities -> direction-free scalars 0
    for nwp in (3, 4, 5):
        cases.append((2, 17, None, np.cumsum(rng.normal(size=(2, nwp)) * 3, 1)))
    for d, N, seq, wseq in cases:
        if wseq is None:
            wd = WD((W(location=col(seq[:, 0]), velocity=col(np.ones(d))), W(location=col(seq[:, -1]), velocity=col(np.ones(d)))))
            pts = seq
        else:
            wd = WD(tuple(W(location=col(wseq[:, k]), velocity=col(np.ones(d))) for k in range(wseq.shape[1])))
            pts = wd.get_waypoint_locations()
        x0 = create_initial_objective_variables(N, pts, wd, d, 3)
        out["initial"].append({"d": d, "N": N, "seq": np.asarray(pts).tolist(), "niw": 0 if wseq is None else wseq.shape[1] - 2,
                               "x0": np.asarray(x0, dtype=float).tolist()})
    for d in (2, 3):
        for _ in range(12):
            p1, p2 = col(rng.normal(size=d) * 6), col(rng.normal(size=d) * 6)
            pad = rng.uniform(1, 4, d)
            if d == 2:
                R, T, Ln = sfcmod.get2DRotationAndTranslationFromPoints(p1, p2)
            else:
                R, T, Ln = sfcmod.get3DRotationAndTranslationFromPoints(p1, p2)
            dims = col(pad).copy(); dims[0, 0] += Ln
            lo, hi = sfcmod.SFC(dims, T, R).getRotatedBounds()
            out["boxes"].append({"d": d, "p1": p1.flatten().tolist(), "p2": p2.flatten().tolist(), "pad": pad.tolist(),
                                 "R": R.tolist(), "lower": lo.flatten().tolist(), "upper": hi.flatten().tolist(), "length": float(Ln)})
    # intervals per corridor chosen from the geometry (SFC_Data.__evaluate_intervals_per_corridor,
    # DS/safe_flight_corridor.py:78-88), incl. ratios that round half to even and a single corridor
    out["intervals"] = []
    seqs = []
    for d in (2, 3):
        for ncorr in (1, 2, 3, 4, 5):
            for _ in range(4):
                seqs.append(np.cumsum(rng.normal(size=(d, ncorr + 1)) * rng.uniform(1, 6), 1))
    seqs.append(np.array([[0.0, 2.0, 7.0, 10.0, 20.0], [0.0, 0.0, 0.0, 0.0, 0.0]]))          # ratios 1, 2.5, 1.5, 5
    seqs.append(np.array([[0.0, 1.0, 4.5, 8.0], [0.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0]]))  # ratios 1, 3.5, 3.5
    for seq in seqs:
        d, ncorr = seq.shape[0], seq.shape[1] - 1
        sfcs = []
        for i in range(ncorr):
            f = sfcmod.get2DRotationAndTranslationFromPoints if d == 2 else sfcmod.get3DRotationAndTranslationFromPoints
            R, T, Ln = f(col(seq[:, i]), col(seq[:, i + 1]))
            sfcs.append(sfcmod.SFC(col([Ln + 2.0] + [2.0] * (d - 1)), T, R))
        for mn in (1, 2):
            data = sfcmod.SFC_Data(tuple(sfcs), seq, mn)
            out["intervals"].append({"d": d, "points": seq.tolist(), "min": mn,
                                     "ipc": [int(v) for v in np.atleast_1d(data.get_intervals_per_corridor())],
                                     "num_intervals": int(data.get_num_intervals())})
    with open(os.path.join(HERE, "build.json"), "w") as f:
        json.dump(out, f)
    print("wrote build.json:", len(out["initial"]), "initial-variable cases,", len(out["boxes"]), "boxes,", len(out["intervals"]), "interval cases")


if __name__ == "__main__":
    main()
