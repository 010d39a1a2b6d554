"""TEST INFRASTRUCTURE: builds and binds tests/hostsim/tg_hostsim.cpp, the single-lane HOST compile of the
device headers (csrc/tg_eval.h, csrc/tg_sqp.h).  It lets the GPU-less container check the exact source that
the CUDA kernels are built from; the product never loads it."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostsim", "tg_hostsim.cpp")
OUT = os.path.join(HERE, "hostsim", "_build_hostsim.so")
CSRC = os.path.join(os.path.dirname(HERE), "trajectory_generator_b200", "csrc")
ND = np.ctypeslib.ndpointer(dtype=np.float64, flags="C")
NI = np.ctypeslib.ndpointer(dtype=np.int32, flags="C")


class HostSim:
    def __init__(self, path):
        self.lib = lib = ctypes.CDLL(path)
        lib.hs_eval.argtypes = [NI, ND, ND, ND, ND, ND, ctypes.c_void_p]
        lib.hs_solve.argtypes = [NI, ND, ND, ctypes.c_int, ctypes.c_double, ctypes.c_int, ND, NI, NI, NI,
                                 ctypes.c_void_p, ctypes.c_int]

    def fd_derivatives(self, pp, x):
        """g[n], J[m, n] as the solver's finite-difference mode forms them at x"""
        L = pp.layout
        g = np.zeros(L.n); J = np.zeros((L.m, L.n))
        self.lib.hs_fd_derivatives.argtypes = [NI, ND, ND, ND, ND, ND, ND]
        xl = np.where(np.isfinite(pp.xl), pp.xl, -np.inf); xu = np.where(np.isfinite(pp.xu), pp.xu, np.inf)
        self.lib.hs_fd_derivatives(pp.spec, pp.par, np.ascontiguousarray(x, dtype=np.float64),
                                   np.ascontiguousarray(xl), np.ascontiguousarray(xu), g, J)
        return g, J

    def eval(self, pp, x, jac=True):
        L = pp.layout
        f = np.zeros(1); g = np.zeros(L.n); c = np.zeros(L.m); J = np.zeros((L.m, L.n))
        self.lib.hs_eval(pp.spec, pp.par, np.ascontiguousarray(x, dtype=np.float64), f, g, c,
                         J.ctypes.data if jac else None)
        return f[0], g, c, J

    def solve(self, pp, maxiter=100, ftol=1e-6, fd=False, trace=False):
        L = pp.layout
        x = np.clip(pp.x0, pp.xl, pp.xu).copy()
        f = np.zeros(1); st = np.zeros(1, np.int32); nit = np.zeros(1, np.int32); nfev = np.zeros(1, np.int32)
        tr = np.zeros((maxiter + 1, L.n + 2)) if trace else None
        self.lib.hs_solve(pp.spec, pp.par, x, maxiter, ftol, 1 if fd else 0, f, st, nit, nfev,
                          tr.ctypes.data if trace else None, tr.size if trace else 0)
        return dict(x=x, f=f[0], status=int(st[0]), nit=int(nit[0]), nfev=int(nfev[0]), trace=tr)


def load():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("tg_eval.h", "tg_sqp.h", "tg_spec.h", "tg_smooth.h")]
    if not os.path.exists(OUT) or any(os.path.getmtime(OUT) < os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-std=c++14", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas",
                        "-DTG_WITH_SQP", "-o", OUT, SRC], check=True, capture_output=True)
    return HostSim(OUT)


def load_variant(tag, flags):
    """A second host build with extra compiler flags (e.g. -DTG_NO_ELIM: the QP stage without the elimination of the
    terminal location rows; -DTG_FUSED_LM_FAR: the lock-step kernels' variant of the stage)."""
    out = os.path.join(HERE, "hostsim", "_build_hostsim_%s.so" % tag)
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("tg_eval.h", "tg_sqp.h", "tg_spec.h", "tg_smooth.h")]
    if not os.path.exists(out) or any(os.path.getmtime(out) < os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-std=c++14", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas",
                        "-DTG_WITH_SQP"] + list(flags) + ["-o", out, SRC], check=True, capture_output=True)
    return HostSim(out)


def scipy_core_solve(hs, pp, maxiter=100, acc=1e-6, record=None):
    """scipy's compiled SLSQP core (scipy.optimize._slsqplib.slsqp, the routine behind the reference's
    minimize(method='SLSQP') call) driven with the host-sim ANALYTIC evaluations, following the loop of
    scipy/optimize/_slsqp_py.py:524-555.  Oracle for the SQP iteration itself."""
    from scipy.optimize._slsqplib import slsqp
    L = pp.layout
    n, m, meq = L.n, L.m, L.meq
    x = np.clip(pp.x0, pp.xl, pp.xu).copy()
    xl = pp.xl.copy(); xu = pp.xu.copy()
    xl[~np.isfinite(xl)] = np.nan; xu[~np.isfinite(xu)] = np.nan
    state = {"acc": acc, "alpha": 0.0, "f0": 0.0, "gs": 0.0, "h1": 0.0, "h2": 0.0, "h3": 0.0, "h4": 0.0, "t": 0.0,
             "t0": 0.0, "tol": 10.0 * acc, "exact": 0, "inconsistent": 0, "reset": 0, "iter": 0,
             "itermax": int(maxiter), "line": 0, "m": m, "meq": meq, "mode": 0, "n": n}
    indices = np.zeros(max(m + 2 * n + 2, 1), dtype=np.int32)
    size = n * (n + 1) // 2 + 3 * m * n - (m + 5 * n + 7) * meq + 9 * m + 8 * n * n + 35 * n + meq * meq + 28
    if m - meq == 0:
        size += 2 * n * (n + 1)
    buffer = np.zeros(max(size, 1)); mult = np.zeros(max(1, m + 2 * n + 2))
    C = np.zeros((max(1, m), n), order="F"); d = np.zeros(max(1, m))
    fx, g, c, J = hs.eval(pp, x)
    C[:m, :] = J; d[:m] = c; g = g.copy()
    while True:
        slsqp(state, fx, g, C, d, x, mult, xl, xu, buffer, indices)
        if state["mode"] == 1:
            fx, _, c, _ = hs.eval(pp, x, jac=False)
            d[:m] = c
        if state["mode"] == -1:
            _, g, _, J = hs.eval(pp, x)
            g = g.copy(); C[:m, :] = J
            if record is not None:
                record.append((state["iter"], fx, x.copy()))
        if abs(state["mode"]) != 1:
            break
    return dict(x=x, f=fx, status=int(state["mode"]), nit=int(state["iter"]))
