"""TEST INFRASTRUCTURE (oracle; never imported by the product): numpy restatement of the reference's spline order
converter, TG/spline_order_converter.py:12-112 (``SmoothingSpline.generate_new_control_points``) -- SURVEY.md 8(f)
row f4, second half.  Pinned against tests/golden/smoothing.json, recorded by running the unmodified reference
(tests/golden/make_golden_smoothing.py).  The product does not implement this row yet: the oracle and its fixtures
are what a CUDA implementation will be checked against.

The problem is an equality-constrained linear least-squares problem per coordinate: control points Q of a spline of
order k with ``int(2.5 * old intervals)`` intervals that minimise sum_t |p_old(t) - p_new(t)|^2 over ``resolution``
samples, with position, velocity and acceleration of the old spline matched at both ends.  The reference hands it
to scipy SLSQP (default ftol 1e-6, finite-difference gradients) from an arc-length initial guess; ``solve_kkt`` is the
exact minimiser SLSQP approaches, ``solve_slsqp`` repeats the reference's call.
"""
from math import factorial

import numpy as np

# TG/matrix_evaluation.py:231-262
_M = {
    1: np.array([[-1, 1], [1, 0]], dtype=float),
    2: 0.5 * np.array([[1, -2, 1], [-2, 2, 1], [1, 0, 0]], dtype=float),
    3: np.array([[-2, 6, -6, 2], [6, -12, 0, 8], [-6, 6, 6, 2], [2, 0, 0, 0]]) / 12,
    4: np.array([[1, -4, 6, -4, 1], [-4, 12, -6, -12, 11], [6, -12, -6, 12, 11], [-4, 4, 6, 4, 1], [1, 0, 0, 0, 0]]) / 24,
    5: np.array([[-1, 5, -10, 10, -5, 1], [5, -20, 20, 20, -50, 26], [-10, 30, 0, -60, 0, 66],
                 [10, -20, -20, 20, 50, 26], [-5, 5, 10, 10, 5, 1], [1, 0, 0, 0, 0, 0]]) / 120,
}


def sampling_matrix(order, num_cont_pts, num_points, r=0, scale=1.0):
    """S [num_cont_pts, num_points] with data = P @ S, for matrix_bspline_evaluation_for_dataset (r = 0,
    TG/matrix_evaluation.py:5-33) and matrix_bspline_derivative_evaluation_for_dataset (r > 0, :104-135): samples
    of np.linspace(0, intervals, num_points) are assigned to interval i by (t >= i) & (t < i + 1), the last interval
    also takes t == i + 1."""
    nint = num_cont_pts - order
    t = np.linspace(0, nint, num_points)
    S = np.zeros((num_cont_pts, num_points))
    M = _M[order]
    marker = 0
    for i in range(nint):
        sel = (t >= i) & (t <= i + 1) if i == nint - 1 else (t >= i) & (t < i + 1)
        steps = t[sel] - i
        L = np.zeros((order + 1, len(steps)))
        for q in range(order - r + 1):
            L[q, :] = steps ** (order - r - q) * (factorial(order - q) / factorial(order - r - q)) / scale ** r
        S[i:i + order + 1, marker:marker + len(steps)] = np.dot(M, L)
        marker += len(steps)
    return S


def initial_control_points(old_pts, num_cont_pts):
    """TG/spline_order_converter.py:83-112: equal arc-length steps along the old control polygon"""
    old_pts = np.asarray(old_pts, dtype=float)
    d, old_n = old_pts.shape
    dist = np.linalg.norm(old_pts[:, 1:] - old_pts[:, :-1], 2, 0)
    for i in range(old_n - 2):
        dist[i + 1] = dist[i + 1] + dist[i]
    nseg = num_cont_pts - 1
    step_len = dist[old_n - 2] / nseg
    seg, cur, prev, step = 0, 0.0, old_pts[:, 0], 0.0
    out = np.zeros((d, num_cont_pts))
    for i in range(nseg):
        v = old_pts[:, seg + 1] - old_pts[:, seg]
        out[:, i] = prev + v / np.linalg.norm(v) * step
        prev = out[:, i]
        step = step_len
        cur = cur + step
        if dist[seg] < cur:
            tmp = dist - cur
            tmp[tmp < 0] = np.inf
            seg = int(np.argmin(tmp))
            step = cur - dist[seg - 1]
            prev = old_pts[:, seg]
    out[:, -1] = old_pts[:, -1]
    return out


class SmoothingProblem:
    def __init__(self, new_order, old_control_points, old_scale_factor, old_order, resolution):
        P = np.asarray(old_control_points, dtype=float)
        self.d, old_n = P.shape
        self.k = int(new_order)
        old_int = old_n - old_order
        self.N = int(old_int * 2.5) + self.k                              # :71-75
        self.scale = old_int * old_scale_factor / (self.N - self.k)       # :77-82
        self.x0 = initial_control_points(P, self.N)
        # objective (:39-46): samples of both splines at `resolution` points
        self.S = sampling_matrix(self.k, self.N, resolution)
        self.Y = P @ sampling_matrix(old_order, old_n, resolution)
        # constraints (:48-69): position, velocity, acceleration at the two end samples
        self.C = np.hstack([sampling_matrix(self.k, self.N, 2, r, self.scale) for r in range(3)])         # [N, 6]
        self.b = np.hstack([P @ sampling_matrix(old_order, old_n, 2, r, old_scale_factor) for r in range(3)])   # [d, 6]

    def objective(self, x):
        Q = np.reshape(x, (self.d, self.N))
        return float(np.sum((self.Y - Q @ self.S) ** 2))

    def constraints(self, x):
        """old - new, in the reference's order: positions, velocities, accelerations, each flattened [d, 2]"""
        Q = np.reshape(x, (self.d, self.N))
        r = self.b - Q @ self.C
        return np.concatenate([r[:, 0:2].flatten(), r[:, 2:4].flatten(), r[:, 4:6].flatten()])

    def solve_kkt(self):
        """exact minimiser: [2 S S', C; C', 0] [q; lam] = [2 S y; b] per coordinate"""
        n = self.N
        K = np.zeros((n + 6, n + 6))
        K[:n, :n] = 2 * self.S @ self.S.T
        K[:n, n:] = self.C
        K[n:, :n] = self.C.T
        Q = np.zeros((self.d, n))
        for c in range(self.d):
            rhs = np.concatenate([2 * self.S @ self.Y[c], self.b[c]])
            Q[c] = np.linalg.solve(K, rhs)[:n]
        return Q

    def solve_slsqp(self):
        """the reference's call (:29-33): scipy SLSQP with default options and finite-difference derivatives"""
        from scipy.optimize import minimize, NonlinearConstraint
        res = minimize(self.objective, x0=self.x0.flatten(), constraints=(NonlinearConstraint(self.constraints, 0, 0)),
                       method="SLSQP")
        return np.reshape(res.x, (self.d, self.N)), res
