"""Generates tests/golden/smoothing.json by running the UNMODIFIED reference's SmoothingSpline
(TG/spline_order_converter.py, imported in place through oracle/ref_import.py).  Build container only:
    python tests/golden/make_golden_smoothing.py"""
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
import ref_import  # noqa: E402

CASES = {
    # test_spline_order_converter.py:9-31: the shipped 2-D cubic, converted to order 4 with 100 samples
    "demo_3_to_4": dict(cp=[[-3, -4, -2, -.5, 1, 0, 2, 3.5, 3], [.5, 3.5, 6, 5.5, 3.7, 2, -1, 2, 5]], scale=1.0, old_order=3,
                        new_order=4, resolution=100),
    "cubic_to_quintic_3d": dict(cp=[[0, 1, 3, 4, 6, 7.5, 9], [0, 2, 1, -1, 0, 2, 3], [1, 1.5, 1, 2, 3, 2.5, 2]], scale=0.7,
                                old_order=3, new_order=5, resolution=80),
    "cubic_to_cubic": dict(cp=[[0, 1, 2.5, 4, 5, 7], [0, 1.5, 0.5, -1, 0.5, 1]], scale=1.3, old_order=3, new_order=3,
                           resolution=60),
}


def main():
    ref_import.namespace()
    with contextlib.redirect_stdout(io.StringIO()):
        soc = importlib.import_module("trajectory_generation.spline_order_converter")
    out = {"cases": {}}
    for name, c in CASES.items():
        cp = np.array(c["cp"], dtype=float)
        sm = soc.SmoothingSpline(c["new_order"], cp.shape[0], c["resolution"])
        captured = {}
        real = soc.minimize

        def spy(*a, **k):
            captured["res"] = real(*a, **k)
            return captured["res"]
        soc.minimize = spy
        try:
            with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                new_cp, new_scale = sm.generate_new_control_points(cp, c["scale"], c["old_order"])
        finally:
            soc.minimize = real
        res = captured["res"]
        x0 = sm.create_initial_control_points(cp, c["old_order"], new_cp.shape[1])
        out["cases"][name] = dict(c, new_control_points=np.asarray(new_cp).tolist(), new_scale_factor=float(new_scale),
                                  initial_control_points=np.asarray(x0).tolist(), status=int(res.status), nit=int(res.nit),
                                  fun=float(res.fun))
        print("%-22s N %d -> %d  scale %.4f  status %d nit %d fun %.3e" % (name, cp.shape[1], new_cp.shape[1], new_scale,
                                                                         res.status, res.nit, res.fun))
    with open(os.path.join(HERE, "smoothing.json"), "w") as f:
        json.dump(out, f)
    print("wrote", os.path.join(HERE, "smoothing.json"))


if __name__ == "__main__":
    main()
