"""PathGenerator: the path problem written as a shape descriptor + parameter row for the
trajectory kernels, against fixtures recorded from the UNMODIFIED reference's PathGenerator
(tests/golden/make_golden_path.py).  CPU side: packing, the SLSQP-facing values and full solves through the
single-lane host build of the kernel source."""
import numpy as np
import pytest

import helpers
import path_problems


def _packed(name):
    from trajectory_generator_b200.problem import pack_problem
    d, cc, kw = path_problems.ALL[name](helpers.product_namespace())
    pp = pack_problem(d, cc, kw.get("objective_function_type", "minimal_velocity_path"),
                      kw.get("num_intervals_free_space"), path_mode="indirect" if kw.get("isIndirect") else "direct")
    return pp


@pytest.mark.parametrize("name", list(path_problems.ALL))
def test_path_problem_is_what_the_reference_hands_to_slsqp(native_lib, hostsim, name):
    """sizes, x0, bounds; objective and SLSQP-ordered constraint values at x0 and at a perturbed point (1e-9)"""
    G = helpers.load_golden("path_generator.json")["problems"][name]
    pp = _packed(name)
    L = pp.layout
    assert (L.d, L.N, L.n, L.m, L.meq) == (G["dimension"], G["N"], G["n"], G["m"], G["meq"])
    assert np.array_equal(np.clip(pp.x0, pp.xl, pp.xu), np.array(G["x0"]))
    assert np.array_equal(pp.xl, np.array(G["xl"])) and np.array_equal(pp.xu, np.array(G["xu"]))
    for xk, fk, ck in (("x0", "f_x0", "c_x0"), ("x_test", "f_test", "c_test")):
        f, g, c, J = hostsim.eval(pp, np.array(G[xk]))
        assert abs(f - G[fk]) <= 1e-9 * max(1.0, abs(G[fk]))
        assert helpers.relerr(c, G[ck]) <= 1e-9, (name, xk)


@pytest.mark.parametrize("name", list(path_problems.ALL))
def test_path_solves_reproduce_the_reference(native_lib, hostsim, name):
    """finite-difference mode (the reference's iterates) under tests/parity_contract.py: the fixtures the unmodified
    reference reproduces from x0 + k ulp agree within 1e-5 with its iteration count; the two that end in a flat valley
    (the reference lands 5e-5 / 8e-4 from itself) are held to the reference's own scatter."""
    import parity_contract
    import tg_oracle
    s = helpers.load_golden("path_generator.json")["problems"][name]["solve"]
    pp = _packed(name)
    L = pp.layout
    mine = hostsim.solve(pp, fd=True)

    def cons(x):
        return hostsim.eval(pp, x, jac=False)[2]
    parity_contract.check(name, s, L.d * L.N, mine["x"], mine["status"], mine["nit"], mine["f"], cons, L.meq)
    assert abs(mine["f"] - s["fun"]) <= 1e-6 * max(1.0, abs(s["fun"]))


def test_path_mode_errors(native_lib):
    from trajectory_generator_b200.problem import pack_problem
    ns = helpers.product_namespace()
    d, cc, kw = path_problems.directions_curvature(ns)
    with pytest.raises(Exception, match="Invalid objective function type"):
        pack_problem(d, cc, "minimal_velocity_and_time_path", path_mode="direct")      # not one of generate_path's three
    with pytest.raises(Exception, match="not supported"):
        pack_problem(d, cc, "minimal_velocity_path", path_mode="indirect")             # indirect + direction
    with pytest.raises(ValueError):
        pack_problem(d, cc, "minimal_velocity_path", path_mode="fast")


def test_alias_module_and_no_device_behaviour(native_lib):
    """the reference's import path works; without a CUDA device the call fails loudly (no CPU path)"""
    import torch
    from trajectory_generation.path_generator import PathGenerator
    d, cc, kw = path_problems.velocities_ignored_obstacle(helpers.product_namespace())
    gen = PathGenerator(d)
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        gen.generate_path(cc, **kw)

