"""Host logic: ConstraintsContainer -> packed problem (descriptor, parameter row, x0, bounds) against the
fixtures recorded from the reference, and the vectorised synthetic generators against the packer."""
import numpy as np
import pytest

import helpers
import problems


@pytest.mark.parametrize("name", list(problems.ALL))
def test_pack_matches_reference_fixture(native_lib, name):
    from trajectory_generator_b200.problem import pack_problem
    G = helpers.load_golden()["problems"][name]
    d, cc, kw = problems.ALL[name](helpers.product_namespace())
    pp = pack_problem(d, cc, kw.get("objective_function_type", "minimal_velocity_and_time_path"),
                      kw.get("num_intervals_free_space"))
    L = pp.layout
    assert (L.d, L.N, L.n, L.m, L.meq) == (G["dimension"], G["N"], G["n"], G["m"], G["meq"])
    assert np.array_equal(np.clip(pp.x0, pp.xl, pp.xu), np.array(G["x0"]))
    assert np.array_equal(pp.xl, np.array(G["xl"])) and np.array_equal(pp.xu, np.array(G["xu"]))


def test_invalid_inputs_raise_like_the_reference(native_lib):
    from trajectory_generator_b200.problem import pack_problem
    ns = helpers.product_namespace()
    d, cc, kw = problems.obstacle2d(ns)
    with pytest.raises(Exception, match="Invalid objective function type"):
        pack_problem(d, cc, "fastest_path")
    # location-only terminal waypoint: the reference dies with IndexError inside scipy
    W, WD = ns["Waypoint"], ns["WaypointData"]
    col = lambda *v: np.array([[float(x)] for x in v])
    cc2 = ns["ConstraintsContainer"](WD((W(location=col(0, 0)), W(location=col(1, 1), velocity=col(1, 0)))))
    with pytest.raises(IndexError):
        pack_problem(2, cc2)
    with pytest.raises(Exception, match="Bound type"):
        ns["TurningBound"](1.0, "yaw")
    with pytest.raises(Exception, match="general max velocity"):
        ns["DerivativeBounds"](max_upward_velocity=1.0)


@pytest.mark.parametrize("name", ["C2", "C3", "C4", "C5a", "C5c"])
def test_synthetic_generators_match_the_packer(native_lib, name):
    from trajectory_generator_b200 import synthetic as syn
    from trajectory_generator_b200.problem import pack_problem
    b = syn.make(name, 257)
    expected_shape = {"C2": (17, 19, 8, 35), "C3": (37, 18, 16, 18), "C4": (34, 209, 15, 71), "C5a": (17, 11, 8, 11),
                      "C5c": (17, 11, 8, 11)}[name]        # (n, m, meq, P) of SURVEY.md 8 table
    L = b.layout
    assert (L.n, L.m, L.meq, L.P) == expected_shape
    for i in (0, 3, 256):
        d, cc, kw = syn.container_for(b, i)
        pp = pack_problem(d, cc, kw.get("objective_function_type", syn.OBJECTIVE[name]), kw.get("num_intervals_free_space"))
        assert np.array_equal(pp.spec, b.spec)
        assert np.abs(pp.par - b.par[i]).max() <= 1e-13
        assert np.abs(pp.x0 - b.x0[i]).max() <= 1e-13
    # same seed -> same batch (the CPU baseline regenerates it in its worker processes)
    assert np.array_equal(syn.make(name, 257).par, b.par)
