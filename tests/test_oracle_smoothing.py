"""Oracle of the spline order converter (second half; TG/spline_order_converter.py) pinned
against fixtures from the unmodified reference.  Test infrastructure only: the product does not implement the row yet."""
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
