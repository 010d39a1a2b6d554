"""The reference's 24 legacy C symbols executed ON THE DEVICE (tg_legacy_kernel<2>, <3>): the reference's own
known-answer vectors (CC/tests/UnitTest*.cpp, lifted into tests/golden/native_kats.json) at the reference's
tolerances, agreement with the plain-C oracle to 1e-9 on those and on random inputs, and the call pattern of the
reference's ctypes wrappers (CF/obstacle_constraints.py:53-79)."""
import ctypes
import json
import os

import numpy as np
import pytest

import helpers
import legacy_abi

pytestmark = pytest.mark.gpu

KATS = json.load(open(os.path.join(helpers.TESTS, "golden", "native_kats.json")))["kats"]
ORACLE_SO = os.path.join(helpers.ROOT, "oracle", "_build", "libtg_oracle.so")


@pytest.mark.parametrize("i", range(len(KATS)))
def test_reference_known_answers_on_the_device(native_lib, oracle_built, i):
    """includes a15 (find_min_velocity_of_bez_vel_cont_pts_{2,3}: MDM hull min-norm, CC/src/MDMAlgorithmClass.cpp)
    and the cusp case that answers DBL_MAX (UnitTestCrossTermBounds.cpp:116-126)"""
    k = KATS[i]
    got = legacy_abi.call(native_lib, k)
    assert not np.any(np.isnan(got)), (k["fn"], got)
    assert np.allclose(got, k["expect"], rtol=0, atol=k["tol"]), (k["fn"], k["src"], got, k["expect"])
    ref = legacy_abi.call(ctypes.CDLL(ORACLE_SO), k)
    tol = 1e-5 if k["fn"] == "find_min_velocity_of_bez_vel_cont_pts" else 1e-9     # MDM stops at its own 1e-6 test
    assert helpers.relerr(got, ref) <= tol, (k["fn"], got, ref)


def test_random_inputs_against_the_c_oracle(native_lib, oracle_built):
    rng = np.random.default_rng(11)
    orc = ctypes.CDLL(ORACLE_SO)
    worst = 0.0
    for trial in range(60):
        D = 2 + trial % 2
        N = int(rng.integers(4, 40))
        pts = (rng.standard_normal((D, N)).cumsum(1) * 2).flatten().tolist()
        al = float(rng.uniform(0.5, 2.0))
        K = int(rng.integers(1, 40))
        ctr = (rng.standard_normal(D) * 5).tolist()
        cases = [dict(fn="get_spline_curvature_bound", D=D, N=N, pts=pts),
                 dict(fn="get_spline_angular_rate_bound", D=D, N=N, pts=pts, alpha=al),
                 dict(fn="get_spline_centripetal_acceleration_bound", D=D, N=N, pts=pts, alpha=al),
                 dict(fn="find_min_velocity_of_spline", D=D, N=N, pts=pts, alpha=al),
                 dict(fn="getObstaclesConstraintsForSpline", D=D, N=N, pts=pts,
                      centers=(rng.standard_normal((D, K)) * 5).flatten().tolist(), radii=rng.uniform(0.3, 2, K).tolist()),
                 dict(fn="getObstacleConstraintsForIntervals", D=D, N=N, pts=pts, radius=0.7, center=ctr),
                 dict(fn="getObstacleConstraintForSpline", D=D, N=N, pts=pts, radius=0.7, center=ctr)]
        for k in cases:
            worst = max(worst, helpers.relerr(legacy_abi.call(native_lib, k), legacy_abi.call(orc, k)))
    assert worst <= 1e-9, worst


def test_call_pattern_of_the_reference_wrappers(native_lib):
    """CF/obstacle_constraints.py:53-79: one handle made once, the restype re-assigned to an ndpointer of the
    expected shape before every call, the result viewed zero-copy.  Arrays returned by earlier calls on the same
    handle stay intact (the reference hands out a fresh array per call; here the handle keeps the last 64)."""
    ND = np.ctypeslib.ndpointer(dtype=np.float64, ndim=1, flags="C")
    lib = native_lib
    lib.ObstacleConstraints_2.restype = ctypes.c_void_p
    obj = ctypes.c_void_p(lib.ObstacleConstraints_2(0))          # the wrappers pass a dummy 0
    rng = np.random.default_rng(3)
    pts = (rng.standard_normal((2, 9)).cumsum(1) * 2)
    views, copies = [], []
    for K in (3, 1, 5, 3):
        centers = rng.standard_normal((2, K)) * 4
        radii = rng.uniform(0.3, 1.0, K)
        lib.getObstaclesConstraintsForSpline_2.argtypes = [ctypes.c_void_p, ND, ND, ctypes.c_int, ND, ctypes.c_int]
        lib.getObstaclesConstraintsForSpline_2.restype = np.ctypeslib.ndpointer(dtype=ctypes.c_double, shape=(K,))
        out = lib.getObstaclesConstraintsForSpline_2(obj, centers.flatten().astype("float64"), radii.astype("float64"), K,
                                                     pts.flatten().astype("float64"), 9)
        assert out.shape == (K,) and not np.any(np.isnan(out))
        views.append(out); copies.append(np.array(out))
        # per-obstacle distances equal the single-obstacle symbol's
        lib.getObstacleConstraintForSpline_2.argtypes = [ctypes.c_void_p, ND, ctypes.c_int, ctypes.c_double, ND]
        lib.getObstacleConstraintForSpline_2.restype = ctypes.c_double
        for i in range(K):
            one = lib.getObstacleConstraintForSpline_2(obj, pts.flatten().astype("float64"), 9, float(radii[i]),
                                                       np.ascontiguousarray(centers[:, i]))
            assert abs(one - out[i]) <= 1e-12
    for v, c in zip(views, copies):
        assert np.array_equal(np.array(v), c)          # earlier views were not overwritten by later calls
    lib.getObstacleConstraintsForIntervals_2.argtypes = [ctypes.c_void_p, ND, ctypes.c_int, ctypes.c_double, ND]
    lib.getObstacleConstraintsForIntervals_2.restype = np.ctypeslib.ndpointer(dtype=ctypes.c_double, shape=(6,))
    per = lib.getObstacleConstraintsForIntervals_2(obj, pts.flatten().astype("float64"), 9, 0.5, np.array([1.0, 2.0]))
    assert per.shape == (6,) and np.isfinite(per).all()


def test_legacy_library_dropped_at_the_reference_path(native_lib, tmp_path):
    """the product library copied to the path the reference's wrappers load
    (<package>/constraint_functions/TrajectoryConstraintsCCode/build/src/libTrajectoryConstraints.so) serves the
    wrappers' calls -- replayed here with the wrappers' own ctypes declarations (CF/turning_constraints.py:17-60,
    CF/min_velocity_evaluator.py:11-40) on a copy loaded from such a path."""
    import shutil
    from trajectory_generator_b200 import _native
    dst = tmp_path / "constraint_functions" / "TrajectoryConstraintsCCode" / "build" / "src"
    dst.mkdir(parents=True)
    shutil.copy(_native.LIB_PATH, dst / "libTrajectoryConstraints.so")
    lib = ctypes.CDLL(str(dst / "libTrajectoryConstraints.so"))
    ND = np.ctypeslib.ndpointer(dtype=np.float64, ndim=1, flags="C")
    k = next(q for q in KATS if q["fn"] == "get_spline_angular_rate_bound")
    lib.CrossTermBounds_2.argtypes = [ctypes.c_void_p]
    lib.CrossTermBounds_2.restype = ctypes.c_void_p
    obj = lib.CrossTermBounds_2(0)
    lib.get_spline_angular_rate_bound_2.argtypes = [ctypes.c_void_p, ND, ctypes.c_int, ctypes.c_double]
    lib.get_spline_angular_rate_bound_2.restype = ctypes.c_double
    cp = np.array(k["pts"], dtype=float).reshape(2, -1)
    got = lib.get_spline_angular_rate_bound_2(obj, cp.flatten().astype("float64"), cp.shape[1], k["alpha"])
    assert abs(got - k["expect"][0]) <= k["tol"]
    k = next(q for q in KATS if q["fn"] == "find_min_velocity_of_spline" and q["D"] == 2 and q["N"] == 6)
    lib.DerivativeBounds_2.argtypes = [ctypes.c_void_p]
    lib.DerivativeBounds_2.restype = ctypes.c_void_p
    obj = lib.DerivativeBounds_2(0)
    lib.find_min_velocity_of_spline_2.argtypes = [ctypes.c_void_p, ND, ctypes.c_int, ctypes.c_double]
    lib.find_min_velocity_of_spline_2.restype = ctypes.c_double
    cp = np.array(k["pts"], dtype=float).reshape(2, -1)
    got = lib.find_min_velocity_of_spline_2(obj, cp.flatten().astype("float64"), cp.shape[1], k["alpha"])
    assert abs(got - k["expect"][0]) <= 1e-12          # the value the reference prints at import time
