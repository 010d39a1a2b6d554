"""CPU checks of the exact source the CUDA kernels are compiled from (single-lane host build, tests/hostsim):
evaluation against the oracle, analytic Jacobians against the Jacobian oracle (4th-order central differences of
the oracle closures), the SQP iteration against scipy's compiled SLSQP core and against the reference's solves."""
import numpy as np
import pytest

import helpers
import hostsim_loader
import problems


def _problem(name):
    from trajectory_generator_b200.problem import pack_problem
    import tg_oracle
    d, cc, kw = problems.ALL[name](helpers.product_namespace())
    obj = kw.get("objective_function_type", "minimal_velocity_and_time_path")
    pp = pack_problem(d, cc, obj, kw.get("num_intervals_free_space"))
    op = tg_oracle.OracleProblem(d, cc, obj, kw.get("num_intervals_free_space"))
    return pp, op


@pytest.mark.parametrize("name", list(problems.ALL))
def test_values_and_jacobians(native_lib, hostsim, name):
    pp, op = _problem(name)
    L = pp.layout
    for seed in (1, 2, 3):
        x = np.clip(problems.test_point(pp.x0, L.d, L.N, seed), pp.xl, pp.xu)
        f, g, c, J = hostsim.eval(pp, x)
        assert abs(f - op.fun(x)) <= 1e-12 * max(1.0, abs(op.fun(x)))
        assert helpers.relerr(c, op.cons(x)) <= 1e-12
        # linear rows: exact constant matrix
        lin = op.linear_jacobian()
        rows = ~np.isnan(lin[:, 0])
        assert np.abs(J[rows] - lin[rows]).max() <= 1e-15 if rows.any() else True
        # nonlinear rows: Jacobian oracle (only meaningful away from kinks; h small enough to stay on a branch)
        e, trusted = op.jacobian_error(J, x)
        assert trusted.mean() >= 0.97, (name, seed, trusted.mean())
        assert e[trusted].max() <= 1e-9, (name, seed, e[trusted].max())
        eg, tg = op.jacobian_error(g[None, :], x, fun=lambda z: np.atleast_1d(op.fun(z)))
        assert tg.all() and eg.max() <= 1e-9


CONVERGING = ["obstacle2d", "obstacles8", "intermediate_curvature", "sfc3d", "sfc3d_four"]


@pytest.mark.parametrize("name", CONVERGING)
def test_sqp_follows_scipy_slsqp_core(native_lib, hostsim, name):
    """Same evaluations (analytic) into scipy's own SLSQP core and into tg_sqp.h: same exit mode, same number of
    major iterations, same iterates."""
    pp, _ = _problem(name)
    rec = []
    ref = hostsim_loader.scipy_core_solve(hostsim, pp, record=rec)
    mine = hostsim.solve(pp, trace=True)
    assert (mine["status"], mine["nit"]) == (ref["status"], ref["nit"]) == (0, ref["nit"])
    assert np.abs(mine["x"] - ref["x"]).max() <= 1e-8
    for it, fx, xk in rec[:10]:
        assert np.abs(mine["trace"][it - 1][2:] - xk).max() <= 1e-8


def test_sqp_augmented_subproblem_path(native_lib, hostsim):
    """c1_sfc2d linearises inconsistently from iteration 2 on: exercises the slack-variable subproblem
    (penalty on E's diagonal as in SLSQP's LSQ).  Iterates follow scipy's core while the factor is well conditioned."""
    pp, _ = _problem("c1_sfc2d")
    rec = []
    ref = hostsim_loader.scipy_core_solve(hostsim, pp, record=rec)
    mine = hostsim.solve(pp, trace=True)
    assert mine["status"] == ref["status"] == 0
    for it, fx, xk in rec[:7]:
        assert np.abs(mine["trace"][it - 1][2:] - xk).max() <= 1e-8
    # from iteration ~11 on the BFGS factor is ill conditioned and the two QP solvers' rounding separates the
    # iterates (28 vs 32 major iterations); both stop at the same optimum within what ftol = 1e-6 resolves
    assert np.abs(mine["x"] - ref["x"]).max() <= 1e-4


@pytest.mark.parametrize("name", list(problems.SOLVE))
def test_fd_mode_reproduces_reference_solves(native_lib, hostsim, name):
    """Finite-difference emulation against EVERY recorded solve of the unmodified reference, under the contract of
    tests/parity_contract.py (read off the reference's own reproducibility runs stored in the fixture)."""
    import parity_contract
    pp, op = _problem(name)
    s = helpers.load_golden()["problems"][name]["solve"]
    mine = hostsim.solve(pp, fd=True)
    parity_contract.check(name, s, pp.layout.ia + 1, mine["x"], mine["status"], mine["nit"], mine["f"], op.cons, op.meq)


@pytest.mark.parametrize("name", ["c1_curvature", "intermediate_waypoints", "unicycle2"])
def test_iteration_limit_flag_matches_reference_analytic_mode(native_lib, hostsim, name):
    s = helpers.load_golden()["problems"][name]["solve"]
    mine = hostsim.solve(pp := _problem(name)[0])
    assert (mine["status"], mine["nit"]) == (s["status"], s["nit"]) == (9, 100)
    del pp


def test_synthetic_sample_against_scipy_core(native_lib, hostsim):
    """First problems of the C3 / C4 batches: every one converges and lands on scipy-core's answer."""
    from trajectory_generator_b200 import synthetic as syn
    from trajectory_generator_b200.problem import pack_problem
    for name, count in (("C3", 4), ("C4", 4), ("C2", 6)):
        b = syn.make(name, 64)
        for i in range(count):
            d, cc, kw = syn.container_for(b, i)
            pp = pack_problem(d, cc, kw.get("objective_function_type", syn.OBJECTIVE[name]), kw.get("num_intervals_free_space"))
            ref = hostsim_loader.scipy_core_solve(hostsim, pp)
            mine = hostsim.solve(pp)
            if ref["status"] == 0:
                assert mine["status"] == 0, (name, i)
                assert np.abs(mine["x"] - ref["x"]).max() <= 1e-5, (name, i)     # north-star tolerance


@pytest.mark.parametrize("name", list(problems.ALL))
def test_fd_mode_derivatives_match_scipy_forward_differences(native_lib, hostsim, name):
    """The solver's finite-difference mode (per-variable sweep of the light blocks + item-parallel sweeps of the
    turning and obstacle rows) against scipy's own forward differences of the reference closures recorded in the
    fixtures (jac_fd_test / grad_fd_test at x_test).  Both difference the same function with the same steps; the
    function values agree to ~1e-15, so the quotients agree to ~1e-15 / 1.5e-8."""
    pp, _ = _problem(name)
    G = helpers.load_golden()["problems"][name]
    x = np.array(G["x_test"])
    g, J = hostsim.fd_derivatives(pp, x)
    Jref = np.array(G["jac_fd_test"]); gref = np.array(G["grad_fd_test"])
    with np.errstate(all="ignore"):
        e = np.abs(J - Jref) / np.maximum(1.0, np.abs(Jref))
    e = np.where(np.isfinite(e), e, 0.0)
    assert e.max() <= 2e-6, (name, e.max(), np.unravel_index(e.argmax(), e.shape))
    assert np.abs(g - gref).max() <= 2e-6 * max(1.0, np.abs(gref).max())


def _ldl_dense(Lm, Dd, n):
    """B = L D L' from the solver's storage (unit lower factor, column i at Lm[i*n + j], j > i)."""
    Lf = np.eye(n)
    for i in range(n):
        for j in range(i + 1, n):
            Lf[j, i] = Lm[i * n + j]
    return Lf @ np.diag(Dd) @ Lf.T


@pytest.mark.parametrize("n", [1, 2, 5, 17, 37])
def test_factor_update_against_dense_arithmetic(hostsim, n):
    """tg_ldl_update (three phases: recurrence / scalars of every column at once / columns of L) against
    B + sigma z z' formed densely, for the two updates of a damped BFGS step (sigma > 0, then sigma < 0 with a
    positive definite result), and for a large positive update that takes the alpha > 4 branch."""
    import ctypes
    ND = np.ctypeslib.ndpointer(dtype=np.float64, flags="C")
    hostsim.lib.hs_ldl_update.argtypes = [ctypes.c_int, ctypes.c_double, ND, ND, ND]
    hostsim.lib.hs_ldl_update.restype = None
    rng = np.random.default_rng(100 + n)
    for trial in range(20):
        Lm = np.zeros(n * n)
        for i in range(n):
            for j in range(i + 1, n):
                Lm[i * n + j] = rng.normal() * 0.5
        Dd = rng.uniform(0.1, 3.0, n)
        B = _ldl_dense(Lm, Dd, n)
        s = rng.normal(size=n)
        u = B @ s + rng.normal(size=n) * 0.3
        if u @ s < 0.2 * (s @ B @ s):
            u = u + s * (1.0 + abs(u @ s)) / (s @ s)
        scale = 50.0 if trial % 4 == 3 else 1.0            # a large update: alpha = t'/t > 4 in some columns
        steps = [(scale / (u @ s), u.copy()), (-1.0 / (s @ B @ s), (B @ s).copy())]
        for sigma, z in steps:
            B = B + sigma * np.outer(z, z)
            hostsim.lib.hs_ldl_update(n, float(sigma), np.ascontiguousarray(z), Lm, Dd)
            got = _ldl_dense(Lm, Dd, n)
            assert np.all(Dd > 0)
            assert np.abs(got - B).max() <= 1e-10 * max(1.0, np.abs(B).max()), (n, trial, sigma)


def test_factor_update_with_zero_weight_is_a_no_op(hostsim):
    import ctypes
    ND = np.ctypeslib.ndpointer(dtype=np.float64, flags="C")
    hostsim.lib.hs_ldl_update.argtypes = [ctypes.c_int, ctypes.c_double, ND, ND, ND]
    hostsim.lib.hs_ldl_update.restype = None
    n = 4
    Lm = np.arange(16, dtype=np.float64) / 10; Dd = np.array([1.0, 2.0, 3.0, 4.0])
    L0, D0 = Lm.copy(), Dd.copy()
    hostsim.lib.hs_ldl_update(n, 0.0, np.ones(n), Lm, Dd)
    assert np.array_equal(Lm, L0) and np.array_equal(Dd, D0)



def test_augmented_subproblem_on_infeasible_corridor_problems(native_lib, hostsim):
    """3-D corridor problems made infeasible (velocity / acceleration bounds far too tight): every iteration
    solves the slack-variable subproblem, and its violation scan takes the corridor rows through the hull points
    with the slack column on top.  The iterates follow scipy's compiled SLSQP core."""
    from trajectory_generator_b200 import synthetic as syn
    from trajectory_generator_b200.problem import pack_problem
    b = syn.make("C4", 8)
    L = b.layout
    for i in range(3):
        d, cc, kw = syn.container_for(b, i)
        pp = pack_problem(d, cc, kw.get("objective_function_type", syn.OBJECTIVE["C4"]), kw.get("num_intervals_free_space"))
        pp.par = pp.par.copy()
        pp.par[L.p_maxv] = 0.05
        pp.par[L.p_maxa] = 1e-4
        rec = []
        ref = hostsim_loader.scipy_core_solve(hostsim, pp, maxiter=12, record=rec)
        mine = hostsim.solve(pp, maxiter=12, trace=True)
        assert (mine["status"], mine["nit"]) == (ref["status"], ref["nit"]) == (9, 12)
        assert len(rec) >= 10
        for it, fx, xk in rec[:10]:
            assert np.abs(mine["trace"][it - 1][2:] - xk).max() <= 1e-8, (i, it)
