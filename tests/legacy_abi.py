"""Shared helper: call one of the reference's 24 C symbols on any library that exports them
(the plain-C oracle, the reference's own C++ build, the CUDA product)."""
import ctypes

import numpy as np

ND = np.ctypeslib.ndpointer(dtype=np.float64, ndim=1, flags="C")
CTORS = {"get_spline_curvature_bound": "CrossTermBounds", "get_spline_angular_rate_bound": "CrossTermBounds",
         "get_spline_centripetal_acceleration_bound": "CrossTermBounds",
         "find_min_velocity_of_spline": "DerivativeBounds",
         "getObstaclesConstraintsForSpline": "ObstacleConstraints",
         "getObstacleConstraintsForIntervals": "ObstacleConstraints",
         "getObstacleConstraintForSpline": "ObstacleConstraints",
         "find_min_velocity_of_bez_vel_cont_pts": "ControlPointDerivativeBounds"}
SYMBOLS = sorted(set("%s_%d" % (n, D) for n in list(CTORS) + list(CTORS.values()) for D in (2, 3)))


def call(lib, k):
    """k: one entry of tests/golden/native_kats.json.  Returns a list of doubles."""
    D, fn, N = k["D"], k["fn"], k["N"]
    ctor = getattr(lib, "%s_%d" % (CTORS[fn], D))
    ctor.restype = ctypes.c_void_p
    h = ctypes.c_void_p(ctor())
    f = getattr(lib, "%s_%d" % (fn, D))
    pts = np.ascontiguousarray(k["pts"], dtype=np.float64)
    if fn == "get_spline_curvature_bound" or fn == "find_min_velocity_of_bez_vel_cont_pts":
        f.argtypes = [ctypes.c_void_p, ND, ctypes.c_int]; f.restype = ctypes.c_double
        return [f(h, pts, N)]
    if fn in ("get_spline_angular_rate_bound", "get_spline_centripetal_acceleration_bound", "find_min_velocity_of_spline"):
        f.argtypes = [ctypes.c_void_p, ND, ctypes.c_int, ctypes.c_double]; f.restype = ctypes.c_double
        return [f(h, pts, N, float(k["alpha"]))]
    if fn == "getObstaclesConstraintsForSpline":
        K = len(k["radii"])
        f.argtypes = [ctypes.c_void_p, ND, ND, ctypes.c_int, ND, ctypes.c_int]
        f.restype = ctypes.POINTER(ctypes.c_double)
        p = f(h, np.ascontiguousarray(k["centers"], dtype=np.float64), np.ascontiguousarray(k["radii"], dtype=np.float64),
              K, pts, N)
        return [p[i] for i in range(K)]
    ctr = np.ascontiguousarray(k["center"], dtype=np.float64)
    if fn == "getObstacleConstraintsForIntervals":
        f.argtypes = [ctypes.c_void_p, ND, ctypes.c_int, ctypes.c_double, ND]
        f.restype = ctypes.POINTER(ctypes.c_double)
        p = f(h, pts, N, float(k["radius"]), ctr)
        return [p[i] for i in range(N - 3)]
    f.argtypes = [ctypes.c_void_p, ND, ctypes.c_int, ctypes.c_double, ND]; f.restype = ctypes.c_double
    return [f(h, pts, N, float(k["radius"]), ctr)]
