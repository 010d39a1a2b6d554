"""The QP stage solves its subproblem without the terminal location rows (csrc/tg_sqp.h, "Eliminated variables":
pinned control points of zero-velocity waypoints, Householder-rotated end triples of plain location rows).  In exact
arithmetic the iterates are those of the full-space subproblem; here the host build with the elimination is run
against the host build without it (-DTG_NO_ELIM), analytic derivatives, and the iterates must agree to rounding over
the first iterations -- including the augmented subproblem while the eliminated coordinates still have to move (the
slack then moves them too), which is forced with a start waypoint outside its corridor."""
import numpy as np
import pytest

import helpers
import hostsim_loader
import problems
from trajectory_generator_b200 import synthetic as syn
from trajectory_generator_b200.problem import pack_problem, Layout


@pytest.fixture(scope="module")
def builds():
    return (hostsim_loader.load_variant("lmfar", ["-DTG_FUSED_LM_FAR"]),          # elimination, lock-step variant of the stage
            hostsim_loader.load(),                                                # elimination, fused variant
            hostsim_loader.load_variant("noelim", ["-DTG_NO_ELIM"]))              # full-space subproblem


class _P:
    pass


def _synthetic(name, count):
    bt = syn.make(name, count)
    L = bt.layout
    for i in range(count):
        pp = _P()
        pp.spec, pp.par, pp.x0, pp.layout = bt.spec, bt.par[i].copy(), bt.x0[i].copy(), L
        pp.xl = np.full(L.n, -np.inf); pp.xu = np.full(L.n, np.inf)
        pp.xl[L.ia:L.it0] = 10e-8
        if L.it0 < L.n:
            pp.xl[L.it0:] = 0; pp.xu[L.it0:] = L.N - 3
        yield pp


def _first_iterates_agree(a, b, pp, iters, tol):
    ra = a.solve(pp, fd=False, maxiter=iters, trace=True)
    rb = b.solve(pp, fd=False, maxiter=iters, trace=True)
    k = min(ra["nit"], rb["nit"])
    assert k >= 1
    scale = max(1.0, np.abs(rb["trace"][:k, 2:]).max())
    assert np.abs(ra["trace"][:k, 2:] - rb["trace"][:k, 2:]).max() <= tol * scale
    return ra, rb


# (C5, the bicycle shape the reference itself is chaotic on, amplifies rounding by 1e3 per iteration: fewer iterations)
@pytest.mark.parametrize("name,iters,tol", [("C2", 6, 1e-11), ("C3", 6, 1e-11), ("C4", 6, 1e-11), ("C5a", 3, 1e-9)])
def test_eliminated_subproblem_gives_the_full_space_iterates(builds, name, iters, tol):
    far, fused, full = builds
    for pp in _synthetic(name, 6):
        _first_iterates_agree(far, full, pp, iters, tol)
        _first_iterates_agree(fused, full, pp, iters, tol)


def test_augmented_subproblem_while_the_pins_move(builds):
    """start waypoint 40 m outside its corridor: the first linearisations are inconsistent, the augmented problem is
    solved with non-zero pin residuals (coupled slack)"""
    far, fused, full = builds
    for pp in _synthetic("C4", 6):
        pp.par[pp.layout.p_start_loc] += 40.0
        for build in (far, fused):
            ra, rb = _first_iterates_agree(build, full, pp, 3, 1e-8)
            assert ra["status"] == rb["status"]


def test_fixture_shapes_with_and_without_elimination(builds):
    """every fixture shape (pins at the start, target form of the end waypoint, both kinds of terminal block, ...): the
    first iterates agree"""
    far, fused, full = builds
    ns = helpers.product_namespace()
    for name in problems.ALL:
        d, cc, kw = problems.ALL[name](ns)
        pp = pack_problem(d, cc, kw.get("objective_function_type", "minimal_velocity_and_time_path"), kw.get("num_intervals_free_space"))
        # (three iterations: features3d amplifies rounding by 1e3 ... 1e4 per iteration -- 2e-14, 3e-11, 3e-10, 5e-6)
        _first_iterates_agree(far, full, pp, 3, 1e-9)
