"""Shared test helpers."""
import json
import os
import sys

import numpy as np

TESTS = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(TESTS)
for p in (ROOT, os.path.join(ROOT, "oracle"), TESTS):
    if p not in sys.path:
        sys.path.insert(0, p)


def dec(o):
    if isinstance(o, str) and o in ("nan", "inf", "-inf"):
        return float(o)
    if isinstance(o, list):
        return [dec(v) for v in o]
    if isinstance(o, dict):
        return {k: dec(v) for k, v in o.items()}
    return o


def load_golden(name="reference_problems.json"):
    with open(os.path.join(TESTS, "golden", name)) as f:
        return dec(json.load(f))


def product_namespace():
    """Public class names bound to this repo's drop-in package."""
    import trajectory_generator_b200.constraint_data_structures as ds
    names = ["Waypoint", "WaypointData", "DerivativeBounds", "TurningBound", "Obstacle", "SFC", "SFC_Data",
             "get2DRotationAndTranslationFromPoints", "get3DRotationAndTranslationFromPoints", "ConstraintsContainer"]
    return {n: getattr(ds, n) for n in names}


def relerr(a, b):
    a = np.asarray(a, dtype=float); b = np.asarray(b, dtype=float)
    with np.errstate(all="ignore"):
        e = np.abs(a - b) / np.maximum(1.0, np.abs(b))
    e = np.where((a == b) | (np.isnan(a) & np.isnan(b)), 0.0, e)
    return float(np.max(e)) if e.size else 0.0
