"""Synthetic problem batches of the BASELINE.json configurations, generated directly as
packed arrays (shape descriptor + one parameter row and one initial variable vector per problem).

Building 10^5..10^6 ``ConstraintsContainer`` objects in Python would dominate every measurement, so the
generators below vectorise what ``pack_problem`` (problem.py) does for one container: same descriptor, same
parameter-row order, same initial guess (TG/objectives/objective_variables.py:27-105).  ``containers_for``
rebuilds individual problems through the public dataclasses so that tests (and the CPU baseline, which runs the
scipy path on the same problems) can check both routes agree.

Seeds: numpy PCG64(20261018 + config index).
"""
import numpy as np

from . import problem as pk
from .constraint_data_structures import (ConstraintsContainer, DerivativeBounds, Obstacle, SFC, SFC_Data, TurningBound,
                                         Waypoint, WaypointData, get3DRotationAndTranslationFromPoints)

SEED0 = 20261018
CONFIGS = ("C2", "C3", "C4", "C5a", "C5c")
FULL_BATCH = {"C2": 65536, "C3": 262144, "C4": 262144, "C5a": 524288, "C5c": 524288}
OBJECTIVE = {"C2": "minimal_velocity_and_time_path", "C3": "minimal_velocity_and_time_path",
             "C4": "minimal_velocity_path", "C5a": "minimal_velocity_and_time_path",
             "C5c": "minimal_velocity_and_time_path"}


class Batch:
    def __init__(self, name, spec, par, x0, raw):
        self.name, self.spec, self.par, self.x0, self.raw = name, spec, par, x0, raw
        self._layout = None

    @property
    def layout(self):
        """Row / parameter offsets of the shape (asks the native library: tg_layout).  Lazy, so that generating a
        batch and rebuilding its containers (the CPU reference arm of bench.py) never loads the CUDA library."""
        if self._layout is None:
            self._layout = pk.Layout(self.spec)
            assert self.par.shape[1] == self._layout.P and self.x0.shape[1] == self._layout.n
        return self._layout

    def __len__(self):
        return self.par.shape[0]

    def take(self, indices):
        """The problems `indices` as a batch of their own (same descriptor; raw fields sliced along the batch axis)."""
        idx = np.asarray(indices, dtype=np.int64)
        B = len(self)
        raw = {k: (v[idx] if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == B else v)
               for k, v in self.raw.items()}
        return Batch(self.name, self.spec, np.ascontiguousarray(self.par[idx]), np.ascontiguousarray(self.x0[idx]), raw)


def _spec(d, N, objective, **flags):
    spec = np.zeros(pk.SP_COUNT, dtype=np.int32)
    spec[pk.SP_DIM], spec[pk.SP_NCP] = d, N
    spec[pk.SP_OBJECTIVE] = pk.OBJECTIVES.index(objective)
    for k, v in flags.items():
        spec[getattr(pk, "SP_" + k)] = v
    return spec


def _line_init(s, g, N):
    """np.linspace(start, end, N) per problem -> [B, d*N] (row-major d x N)."""
    return pk.line_initial_points(s, g, N).reshape(len(s), -1)


_polyline_init = pk.polyline_initial_points


def _heading(theta):
    return np.stack([np.cos(theta), np.sin(theta)], 1)


def make_c2(B, seed=SEED0 + 2):
    """2-D obstacle avoidance, test_obstacle_trajectory_2D.py:16-53 shape with 8 circular obstacles."""
    rng = np.random.Generator(np.random.PCG64(seed))
    s = rng.uniform(-10, 10, (B, 2))
    L = rng.uniform(6, 9, B)
    th = rng.uniform(0, 2 * np.pi, B)
    g = s + L[:, None] * _heading(th)
    v0 = rng.uniform(0.8, 1.2, B)[:, None] * _heading(th + rng.uniform(-np.pi / 4, np.pi / 4, B))
    v1 = rng.uniform(0.8, 1.2, B)[:, None] * _heading(th + rng.uniform(-np.pi / 4, np.pi / 4, B))
    K = 8
    lo = np.minimum(s, g) - 2.0
    hi = np.maximum(s, g) + 2.0
    ctr = rng.uniform(0, 1, (B, K, 2)) * (hi - lo)[:, None, :] + lo[:, None, :]
    rad = rng.uniform(0.3, 1.0, (B, K))
    for _ in range(64):      # resample obstacles too close to a terminal waypoint
        ds = np.linalg.norm(ctr - s[:, None, :], 2, 2)
        dg = np.linalg.norm(ctr - g[:, None, :], 2, 2)
        bad = (ds < rad + 0.75) | (dg < rad + 0.75)
        if not bad.any():
            break
        new = rng.uniform(0, 1, (B, K, 2)) * (hi - lo)[:, None, :] + lo[:, None, :]
        ctr = np.where(bad[:, :, None], new, ctr)
    N = 8
    spec = _spec(2, N, OBJECTIVE["C2"], START_VEL=1, END_VEL=1, DB_MAXV=1, DB_MAXA=1, TURN=2, NOBST=K)
    ones = np.ones((B, 1))
    par = np.concatenate([s, g, v0, v1, 2.0 * ones, 5.0 * ones, 1.8 * ones,
                          ctr[:, :, 0], ctr[:, :, 1], rad], 1)
    x0 = np.concatenate([_line_init(s, g, N), ones], 1)
    raw = dict(start=s, goal=g, v0=v0, v1=v1, centers=ctr, radii=rad, vmax=2.0, amax=5.0, turn=1.8)
    return Batch("C2", spec, np.ascontiguousarray(par), np.ascontiguousarray(x0), raw)


def make_c3(B, seed=SEED0 + 3):
    """2-D intermediate waypoints with velocities (test_intermediate_waypoints.py:14-55 shape), v_max + curvature."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pts = np.empty((B, 2, 4))
    pts[:, :, 0] = rng.uniform(-10, 10, (B, 2))
    head = rng.uniform(0, 2 * np.pi, B)
    for i in range(1, 4):
        if i > 1:
            head = head + rng.uniform(-np.pi / 2, np.pi / 2, B)
        pts[:, :, i] = pts[:, :, i - 1] + rng.uniform(2, 3, B)[:, None] * _heading(head)
    # velocity along the local polyline direction (central difference of the polyline, one-sided at the ends)
    dirs = np.empty((B, 2, 4))
    dirs[:, :, 0] = pts[:, :, 1] - pts[:, :, 0]
    dirs[:, :, 3] = pts[:, :, 3] - pts[:, :, 2]
    dirs[:, :, 1] = pts[:, :, 2] - pts[:, :, 0]
    dirs[:, :, 2] = pts[:, :, 3] - pts[:, :, 1]
    dirs = dirs / np.linalg.norm(dirs, 2, 1)[:, None, :]
    vel = dirs * rng.uniform(0.5, 2.0, (B, 1, 4))
    N = 17
    spec = _spec(2, N, OBJECTIVE["C3"], START_VEL=1, END_VEL=1, NIW=2, IW_VEL=1, DB_MAXV=1, TURN=1)
    ones = np.ones((B, 1))
    iwl = pts[:, :, 1:3].reshape(B, 4)        # (d, niw).flatten()
    iwv = vel[:, :, 1:3].reshape(B, 4)
    par = np.concatenate([pts[:, :, 0], pts[:, :, 3], vel[:, :, 0], vel[:, :, 3], iwl, iwv, 5.0 * ones, 2.0 * ones], 1)
    cps = _polyline_init(pts, N)
    cum = np.cumsum(np.linalg.norm(pts[:, :, 1:] - pts[:, :, :-1], 2, 1), 1)
    times = (cum / cum[:, -1:])[:, :-1] * (N - 3)
    x0 = np.concatenate([cps.reshape(B, -1), ones, times], 1)
    raw = dict(points=pts, velocities=vel, vmax=5.0, turn=2.0)
    return Batch("C3", spec, np.ascontiguousarray(par), np.ascontiguousarray(x0), raw)


def _rot3(p1, p2):
    """Batched get3DRotationAndTranslationFromPoints (DS/safe_flight_corridor.py:123-146)."""
    delta = p2 - p1
    theta = np.arctan2(delta[:, 2], delta[:, 0])
    c, s = np.cos(theta), np.sin(theta)
    B = len(p1)
    Ry = np.zeros((B, 3, 3))
    Ry[:, 0, 0] = c; Ry[:, 0, 2] = s; Ry[:, 1, 1] = 1; Ry[:, 2, 0] = -s; Ry[:, 2, 2] = c
    v = np.einsum("bij,bj->bi", Ry, delta)
    psi = np.arctan2(v[:, 1], v[:, 0])
    c, s = np.cos(psi), np.sin(psi)
    Rz = np.zeros((B, 3, 3))
    Rz[:, 0, 0] = c; Rz[:, 0, 1] = -s; Rz[:, 1, 0] = s; Rz[:, 1, 1] = c; Rz[:, 2, 2] = 1
    R = np.einsum("bji,bjk->bik", Ry, Rz)         # Ry^T Rz
    T = np.einsum("bji,bj->bi", R, (p1 + p2) / 2)  # R^T mid
    return R, T, np.linalg.norm(delta, 2, 1)


def make_c4(B, seed=SEED0 + 4):
    """3-D safe-flight corridors (test_sfc_trajectory_3D.py:19-89 shape), 4 boxes, ipc = [2,2,2,2]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pts = np.empty((B, 3, 5))
    pts[:, :, 0] = rng.uniform(-10, 10, (B, 3))
    ell = rng.uniform(6, 10, B)
    d = rng.standard_normal((B, 3))
    d /= np.linalg.norm(d, 2, 1)[:, None]
    for i in range(1, 5):
        if i > 1:
            # turn by at most 60 degrees: mix the old direction with a random unit vector
            ang = rng.uniform(0, np.pi / 3, B)
            r = rng.standard_normal((B, 3))
            r -= np.sum(r * d, 1)[:, None] * d
            r /= np.linalg.norm(r, 2, 1)[:, None]
            d = np.cos(ang)[:, None] * d + np.sin(ang)[:, None] * r
        pts[:, :, i] = pts[:, :, i - 1] + (ell * rng.uniform(1.0, 1.45, B))[:, None] * d
    N = 11
    spec = _spec(3, N, OBJECTIVE["C4"], START_VEL=1, END_KIND=1, DB_MAXV=1, DB_MAXA=1, NCORR=4)
    spec[pk.SP_IPC0:pk.SP_IPC0 + 4] = 2
    ones = np.ones((B, 1))
    v0 = pts[:, :, 1] - pts[:, :, 0]
    v0 /= np.linalg.norm(v0, 2, 1)[:, None]
    blocks = [pts[:, :, 0], pts[:, :, 4], v0, 5.0 * ones, 0.3 * ones]
    dims_all = []
    for i in range(4):
        R, T, Ln = _rot3(pts[:, :, i], pts[:, :, i + 1])
        dims = np.stack([Ln + rng.uniform(2, 3, B), rng.uniform(2, 3, B), rng.uniform(2, 4, B)], 1)
        dims_all.append(dims)
        blocks += [np.transpose(R, (0, 2, 1)).reshape(B, 9), T - dims / 2, T + dims / 2]
    par = np.concatenate(blocks, 1)
    cps = _polyline_init(pts, N)
    x0 = np.concatenate([cps.reshape(B, -1), ones], 1)
    raw = dict(points=pts, v0=v0, dims=np.stack(dims_all, 1), vmax=5.0, amax=0.3)
    return Batch("C4", spec, np.ascontiguousarray(par), np.ascontiguousarray(x0), raw)


def make_c5(B, turn_kind="angular_rate", seed=SEED0 + 5):
    """Bicycle / unicycle kinematic trajectories (bicycle_trajectory_3.py / unicycle_trajectory_2.py shape)."""
    rng = np.random.Generator(np.random.PCG64(seed + (0 if turn_kind == "angular_rate" else 1)))
    s = np.array([-5.0, 0.0]) + rng.uniform(-1, 1, (B, 2))
    g = np.array([5.0, 0.0]) + rng.uniform(-1, 1, (B, 2))
    v0 = np.stack([np.zeros(B), rng.uniform(20, 28, B)], 1)
    v1 = np.stack([np.zeros(B), rng.uniform(20, 28, B)], 1)
    turn = rng.uniform(12, 18, B) if turn_kind == "angular_rate" else np.full(B, 8.0 / 28.0)
    N = 8
    name = "C5a" if turn_kind == "angular_rate" else "C5c"
    spec = _spec(2, N, OBJECTIVE[name], START_VEL=1, END_VEL=1, DB_MAXV=1, DB_MAXA=1,
                 TURN=pk.TURN_KINDS[turn_kind])
    ones = np.ones((B, 1))
    par = np.concatenate([s, g, v0, v1, 30.0 * ones, 100.0 * ones, turn[:, None]], 1)
    x0 = np.concatenate([_line_init(s, g, N), ones], 1)
    raw = dict(start=s, goal=g, v0=v0, v1=v1, vmax=30.0, amax=100.0, turn=turn, turn_kind=turn_kind)
    return Batch(name, spec, np.ascontiguousarray(par), np.ascontiguousarray(x0), raw)


def make(name, B=None):
    B = FULL_BATCH[name] if B is None else B
    if name == "C2": return make_c2(B)
    if name == "C3": return make_c3(B)
    if name == "C4": return make_c4(B)
    if name == "C5a": return make_c5(B, "angular_rate")
    if name == "C5c": return make_c5(B, "curvature")
    raise KeyError(name)


def _col(v):
    return np.asarray(v, dtype=np.float64).reshape(-1, 1)


def container_for(batch, i):
    """Problem i of a batch through the public dataclasses: (dimension, ConstraintsContainer, generate kwargs)."""
    r = batch.raw
    name = batch.name
    if name == "C2":
        wd = WaypointData((Waypoint(location=_col(r["start"][i]), velocity=_col(r["v0"][i])),
                           Waypoint(location=_col(r["goal"][i]), velocity=_col(r["v1"][i]))))
        obs = [Obstacle(center=_col(r["centers"][i, k]), radius=float(r["radii"][i, k])) for k in range(r["radii"].shape[1])]
        cc = ConstraintsContainer(wd, DerivativeBounds(r["vmax"], r["amax"]), TurningBound(r["turn"], "angular_rate"),
                                  None, obs)
        return 2, cc, dict()
    if name == "C3":
        wps = tuple(Waypoint(location=_col(r["points"][i, :, k]), velocity=_col(r["velocities"][i, :, k])) for k in range(4))
        cc = ConstraintsContainer(WaypointData(wps), DerivativeBounds(r["vmax"], None), TurningBound(r["turn"], "curvature"))
        return 2, cc, dict(num_intervals_free_space=14)
    if name == "C4":
        pts = [_col(r["points"][i, :, k]) for k in range(5)]
        sfcs = []
        for k in range(4):
            R, T, Ln = get3DRotationAndTranslationFromPoints(pts[k], pts[k + 1])
            dims = r["dims"][i, k]
            sfcs.append(SFC(_col(dims), T, R))
        sfc = SFC_Data(tuple(sfcs), np.concatenate(pts, 1), 1)
        wd = WaypointData((Waypoint(location=pts[0], velocity=_col(r["v0"][i])),
                           Waypoint(location=pts[4], velocity=_col([0, 0, 0]))))
        cc = ConstraintsContainer(wd, DerivativeBounds(r["vmax"], r["amax"]), None, sfc, None)
        return 3, cc, dict(objective_function_type="minimal_velocity_path")
    wd = WaypointData((Waypoint(location=_col(r["start"][i]), velocity=_col(r["v0"][i])),
                       Waypoint(location=_col(r["goal"][i]), velocity=_col(r["v1"][i]))))
    cc = ConstraintsContainer(wd, DerivativeBounds(r["vmax"], r["amax"]), TurningBound(float(r["turn"][i]), r["turn_kind"]))
    return 2, cc, dict(num_intervals_free_space=5)


def c1_problem():
    """Config C1 (BASELINE configs[0]): the literal single problem of test_2D_trajectory.py:24-80 -- 2-D, three
    corridors, start / end waypoints with velocity, v_max 30, a_max 100, `minimal_time_path`,
    num_intervals_free_space = 10.  Returns (dimension, ConstraintsContainer, generate kwargs)."""
    from .constraint_data_structures import get2DRotationAndTranslationFromPoints
    pts = [_col(p) for p in ((-5, 0), (0, 5), (0, -5), (5, 0))]
    dims = [(3, 2), (2, 3), (3, 2)]
    sfcs = []
    for i in range(3):
        R, T, Ln = get2DRotationAndTranslationFromPoints(pts[i], pts[i + 1])
        sfcs.append(SFC(np.array([[Ln + dims[i][0]], [dims[i][1]]]), T, R))
    sfc = SFC_Data(tuple(sfcs), np.concatenate(pts, 1), 1, intervals_per_corridor=np.array([1, 1, 1]))
    wd = WaypointData((Waypoint(location=_col((-5, 0)), velocity=_col((0, 15))),
                       Waypoint(location=_col((5, 0)), velocity=_col((0, 10)))))
    cc = ConstraintsContainer(waypoint_constraints=wd, derivative_constraints=DerivativeBounds(30, 100),
                              turning_constraint=None, sfc_constraints=sfc, obstacle_constraints=None)
    return 2, cc, dict(objective_function_type="minimal_time_path", num_intervals_free_space=10)


def evaluation_points(batch, seed=99):
    """M1 evaluation points x0 + 0.3 N(0,1) on the control-point entries."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = batch.x0.copy()
    k = batch.layout.d * batch.layout.N
    x[:, :k] += 0.3 * rng.standard_normal((len(batch), k))
    return x
