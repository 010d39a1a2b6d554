"""The solve-parity contract of the recorded reference solves (tests/golden/reference_problems.json), shared by
the host-build test (tests/test_hostsim.py) and its GPU twin (tests/test_gpu_first.py).

The reference differentiates by forward differences with h = 1.5e-8, which amplifies last-place differences of
its closures by 1/h = 6.7e7; every fixture therefore records the UNMODIFIED reference against itself from
x0 + k ulp (k = +/-1 .. +/-4: `neighbours`) and the contract is read off those runs, never off the problem's name:

  * `stable` (every reference run ends with status 0 and within 1e-5 of the recorded control points):
        same status, control points and scale factor within 1e-5 (the north-star tolerance), and an iteration
        count inside the range of the reference's own nine runs widened by NIT_SLACK (the reference itself needs
        27 .. 32 iterations on the shipped C1 problem depending on the last place of x0);
  * otherwise, `status_stable` (every reference run ends with the same status): same status; if that status is 0
        the objective lies inside the reference's own range of final objectives (widened by the width of that
        range, nine samples, plus 1e-6 relative = ftol),
        the solution is no further from the recorded one than twice the largest distance among the reference's own
        runs, and it is feasible to the level the reference's own solutions are (never looser than 1e-6 + theirs);
  * otherwise (the reference's status flag itself flips under 1 ulp): the status is one the reference produced,
        with the same feasibility check when it is 0.
  Iteration-limit exits (status 9) also agree on the iteration count.
"""
import numpy as np

NIT_SLACK = 3        # iterations outside the range of the reference's own nine runs that a stable fixture may take
SCATTER_FACTOR = 2   # unstable fixtures: distance to the recorded answer <= this x the largest distance among the reference's own runs


def check(name, golden_solve, ncp, x, status, nit, f, cons, meq):
    """x, status, nit, f: this repo's solve; cons(x): SLSQP-ordered constraint values (oracle); ncp = d*N + 1."""
    s = golden_solve
    nb = s["neighbours"]
    runs = [dict(status=s["status"], nit=s["nit"], fun=s["fun"], dcp=0.0, c_min_ineq=s["c_min_ineq"],
                 c_max_eq=s["c_max_eq"])] + list(nb)
    statuses = sorted(set(r["status"] for r in runs))
    dcp = float(np.abs(np.asarray(x)[:ncp] - np.asarray(s["x"])[:ncp]).max())
    ctx = (name, "status", status, "nit", nit, "dcp", dcp, "f", f)
    if s["status_stable"]:
        assert status == s["status"], ctx
    else:
        assert status in statuses, ctx
    if status == 9:
        assert nit == s["nit"] == 100, ctx
    if s["stable"]:
        assert dcp <= 1e-5, ctx
        assert min(r["nit"] for r in runs) - NIT_SLACK <= nit <= max(r["nit"] for r in runs) + NIT_SLACK, ctx
        return dcp
    if status == 0:
        conv = [r for r in runs if r["status"] == 0]
        # the reference's own final objectives (all nine runs) span [flo, fhi]; this solve may sit outside that
        # sample of nine by no more than its width
        flo, fhi = min(r["fun"] for r in runs), max(r["fun"] for r in runs)
        tol = (fhi - flo) + 1e-6 * max(abs(flo), abs(fhi), 1.0)
        assert flo - tol <= f <= fhi + tol, ctx + (flo, fhi)
        if s["status"] == 0:
            assert dcp <= max(1e-5, SCATTER_FACTOR * max(r["dcp"] for r in conv)), ctx
        c = np.asarray(cons(np.asarray(x)))
        eq_ref = max([r["c_max_eq"] for r in conv] + [0.0])
        in_ref = min([r["c_min_ineq"] for r in conv] + [0.0])
        if meq:
            assert np.abs(c[:meq]).max() <= 1e-6 + eq_ref, ctx
        if len(c) > meq:
            assert c[meq:].min() >= -1e-6 + in_ref, ctx
    return dcp
