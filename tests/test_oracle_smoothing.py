"""Oracle of the spline order converter (second half; TG/spline_order_converter.py) pinned
against fixtures from the unmodified reference, the product's solver (csrc/tg_smooth.h) through the single-lane host
build against the same fixtures, and -- on the GPU -- the drop-in SmoothingSpline class."""
import numpy as np
import pytest

import helpers
import tg_oracle_smoothing as osm

CASES = ["demo_3_to_4", "cubic_to_quintic_3d", "cubic_to_cubic"]


@pytest.mark.parametrize("name", CASES)
def test_smoothing_oracle_matches_reference_fixture(name):
    G = helpers.load_golden("smoothing.json")["cases"][name]
    prob = osm.SmoothingProblem(G["new_order"], np.array(G["cp"], dtype=float), G["scale"], G["old_order"], G["resolution"])
    ref = np.array(G["new_control_points"])
    assert ref.shape == (prob.d, prob.N)
    assert abs(prob.scale - G["new_scale_factor"]) <= 1e-15 * max(1.0, G["new_scale_factor"])
    assert np.array_equal(prob.x0, np.array(G["initial_control_points"]))          # arc-length initial guess, bit for bit
    # the reference's own objective value at its answer, from the restated sampling matrices
    assert abs(prob.objective(ref.flatten()) - G["fun"]) <= 1e-12 * max(1.0, G["fun"])
    # the same scipy call on the restated closures lands on the reference's answer
    Q, res = prob.solve_slsqp()
    assert res.status == G["status"] == 0 and abs(res.nit - G["nit"]) <= 2
    # last-place differences of the closures are amplified by the finite differences: same iteration count -> same
    # answer to 1e-6; one iteration more or less on this flat least-squares valley moves the control points by ~1e-3
    # while the objective agrees to 1e-6
    assert np.abs(Q - ref).max() <= (1e-6 if res.nit == G["nit"] else 5e-3)
    assert abs(res.fun - G["fun"]) <= 1e-6
    # and the reference's answer is the exact constrained least-squares solution up to what ftol = 1e-6 resolves
    K = prob.solve_kkt()
    assert np.abs(prob.constraints(K.flatten())).max() <= 1e-9
    assert prob.objective(K.flatten()) <= G["fun"] + 1e-12
    assert np.abs(K - ref).max() <= 5e-2 and abs(prob.objective(K.flatten()) - G["fun"]) <= 1e-4


def test_sampling_matrix_reproduces_the_cubic_sampler():
    """order 3: the general sampling matrix equals the cubic oracle that is pinned bit for bit in test_sampling.py"""
    import tg_oracle_sampling as osamp
    rng = np.random.default_rng(3)
    P = rng.normal(size=(3, 9))
    assert np.abs(P @ osm.sampling_matrix(3, 9, 50) - osamp.dataset(P, 50)).max() <= 1e-13
    assert np.abs(P @ osm.sampling_matrix(3, 9, 50, 1, 0.7) - osamp.derivative_dataset(P, 1, 0.7, 50)).max() <= 1e-12
    assert np.abs(P @ osm.sampling_matrix(3, 9, 50, 2, 0.7) - osamp.derivative_dataset(P, 2, 0.7, 50)).max() <= 1e-12


def _host_solve(hostsim, G):
    import ctypes
    ND = np.ctypeslib.ndpointer(dtype=np.float64, flags="C")
    lib = hostsim.lib
    lib.hs_smooth_solve.argtypes = [ctypes.c_int] * 4 + [ctypes.c_double, ND, ND, ND, np.ctypeslib.ndpointer(dtype=np.int32, flags="C")]
    lib.hs_smooth_initial.argtypes = [ctypes.c_int, ND, ctypes.c_int, ctypes.c_int, ND]
    cp = np.array(G["cp"], dtype=float)
    prob = osm.SmoothingProblem(G["new_order"], cp, G["scale"], G["old_order"], G["resolution"])
    x0 = np.zeros((prob.d, prob.N))
    lib.hs_smooth_initial(prob.d, np.ascontiguousarray(cp), cp.shape[1], prob.N, x0)
    par = np.concatenate([prob.Y.flatten(), prob.b.flatten()])
    x = x0.flatten().copy(); f = np.zeros(1); nit = np.zeros(1, np.int32)
    status = lib.hs_smooth_solve(prob.d, prob.N, prob.k, G["resolution"], prob.scale, par, x, f, nit)
    return prob, x0, x.reshape(prob.d, prob.N), int(status), int(nit[0]), float(f[0])


def _check_against_reference(prob, G, Q, status, nit, fun):
    """status 0, iteration count within 2 of the reference's, objective within 1e-6 (= ftol), end-point rows satisfied,
    and control points as close to the reference's answer as the data supports: SLSQP stops at ftol = 1e-6 on a flat
    least-squares valley, |K - ref| away from the exact constrained minimiser K (3e-4, 1.7e-2 and 1.4e-4 on the three
    fixtures), and where exactly it stops depends on the last place of its iterates -- a solve may land no further from
    the reference's stopping point than a third of that distance (never asked to be closer than 1e-6)."""
    ref = np.array(G["new_control_points"])
    assert status == G["status"] == 0 and abs(nit - G["nit"]) <= 2
    tol = max(1e-6, 0.35 * np.abs(prob.solve_kkt() - ref).max())
    assert np.abs(Q - ref).max() <= tol, (np.abs(Q - ref).max(), tol)
    assert abs(fun - G["fun"]) <= 1e-6
    assert abs(prob.objective(Q.flatten()) - fun) <= 1e-12 * max(1.0, fun)
    assert np.abs(prob.constraints(Q.flatten())).max() <= 1e-9


@pytest.mark.parametrize("name", CASES)
def test_product_solver_on_the_host_build(native_lib, hostsim, name):
    G = helpers.load_golden("smoothing.json")["cases"][name]
    prob, x0, Q, status, nit, fun = _host_solve(hostsim, G)
    assert np.array_equal(x0, np.array(G["initial_control_points"]))          # arc-length walk, bit for bit
    _check_against_reference(prob, G, Q, status, nit, fun)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_smoothing_spline_on_the_gpu(native_lib, name):
    """the reference's call through the alias module: SmoothingSpline(order, d, resolution).generate_new_control_points"""
    from trajectory_generation.spline_order_converter import SmoothingSpline
    G = helpers.load_golden("smoothing.json")["cases"][name]
    cp = np.array(G["cp"], dtype=float)
    prob = osm.SmoothingProblem(G["new_order"], cp, G["scale"], G["old_order"], G["resolution"])
    sm = SmoothingSpline(G["new_order"], cp.shape[0], G["resolution"])
    x0 = sm.create_initial_control_points(cp, G["old_order"], prob.N)
    assert np.abs(x0 - np.array(G["initial_control_points"])).max() <= 1e-14
    Q, scale = sm.generate_new_control_points(cp, G["scale"], G["old_order"])
    assert Q.shape == (prob.d, prob.N) and abs(scale - G["new_scale_factor"]) <= 1e-15 * max(1.0, G["new_scale_factor"])
    r = sm.last_result
    _check_against_reference(prob, G, Q, int(r["status"][0]), int(r["nit"][0]), float(r["fun"][0]))


@pytest.mark.gpu
def test_smoothing_batch_equals_one_at_a_time(native_lib):
    from trajectory_generator_b200.spline_order_converter import SmoothingSpline
    rng = np.random.default_rng(8)
    base = np.array(helpers.load_golden("smoothing.json")["cases"]["demo_3_to_4"]["cp"], dtype=float)
    olds = base[None] + 0.2 * rng.normal(size=(9,) + base.shape)
    sm = SmoothingSpline(4, 2, 100)
    Q, scale = sm.generate_new_control_points_batch(olds, [1.0] * 9, 3)
    assert (sm.last_result["status"] == 0).all()
    for i in (0, 4, 8):
        one, s1 = sm.generate_new_control_points(olds[i], 1.0, 3)
        assert np.array_equal(one, Q[i]) and s1 == scale[i]
    # 3 x 30 = 90 variables (more than a 64-lane pass): solved; end-point rows satisfied
    old = np.cumsum(rng.normal(size=(3, 14)), 1)
    sm3 = SmoothingSpline(3, 3, 50)
    Q3, s3 = sm3.generate_new_control_points(old, 1.0, 3)
    prob = osm.SmoothingProblem(3, old, 1.0, 3, 50)
    assert Q3.shape == (3, 30) and int(sm3.last_result["status"][0]) in (0, 9)
    assert np.abs(prob.constraints(Q3.flatten())).max() <= 1e-6
    assert prob.objective(Q3.flatten()) <= prob.objective(prob.x0.flatten())
    with pytest.raises(RuntimeError, match="unsupported shape"):
        SmoothingSpline(3, 3, 50).generate_new_control_points(np.cumsum(rng.normal(size=(3, 30)), 1), 1.0, 3)      # 3 x 70 variables
