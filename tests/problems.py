"""Problem definitions shared by the golden-fixture generator (which builds them
with the REFERENCE's classes) and the tests (which build them with this repo's
drop-in classes).  Each function takes a namespace ``ns`` holding the public
class names and returns ``(dimension, container, generate_kwargs)``.

Sources: the reference's demo scripts, cited per problem.
"""
import numpy as np


def _col(*v):
    return np.array([[float(x)] for x in v])


def c1_sfc2d(ns):
    """test_2D_trajectory.py:24-80 as shipped (config C1)."""
    W, WD, DB = ns["Waypoint"], ns["WaypointData"], ns["DerivativeBounds"]
    pts = [_col(-5, 0), _col(0, 5), _col(0, -5), _col(5, 0)]
    dims = [(3, 2), (2, 3), (3, 2)]
    sfcs = []
    for i in range(3):
        R, T, L = ns["get2DRotationAndTranslationFromPoints"](pts[i], pts[i + 1])
        sfcs.append(ns["SFC"](np.array([[L + dims[i][0]], [dims[i][1]]]), T, R))
    sfc = ns["SFC_Data"](tuple(sfcs), np.concatenate(pts, 1), 1, intervals_per_corridor=np.array([1, 1, 1]))
    wd = WD((W(location=_col(-5, 0), velocity=_col(0, 15)), W(location=_col(5, 0), velocity=_col(0, 10))))
    cc = ns["ConstraintsContainer"](waypoint_constraints=wd, derivative_constraints=DB(30, 100),
                                    turning_constraint=None, sfc_constraints=sfc, obstacle_constraints=None)
    return 2, cc, dict(objective_function_type="minimal_time_path", num_intervals_free_space=10)


def c1_curvature(ns):
    """test_2D_trajectory.py with the commented curvature bound (:54-55), SFC off (config C1')."""
    W, WD, DB, TB = ns["Waypoint"], ns["WaypointData"], ns["DerivativeBounds"], ns["TurningBound"]
    wd = WD((W(location=_col(-5, 0), velocity=_col(0, 15)), W(location=_col(5, 0), velocity=_col(0, 10))))
    cc = ns["ConstraintsContainer"](waypoint_constraints=wd, derivative_constraints=DB(30, 100),
                                    turning_constraint=TB(1, "curvature"))
    return 2, cc, dict(objective_function_type="minimal_time_path", num_intervals_free_space=10)


def obstacle2d(ns):
    """test_obstacle_trajectory_2D.py:16-53 (config C2 shape with its single obstacle)."""
    W, WD, DB, TB, Ob = ns["Waypoint"], ns["WaypointData"], ns["DerivativeBounds"], ns["TurningBound"], ns["Obstacle"]
    w1 = W(location=_col(3, 4)); w2 = W(location=_col(7, 10))
    w1.velocity = _col(1, 0); w2.velocity = _col(1, 1)
    cc = ns["ConstraintsContainer"](waypoint_constraints=WD((w1, w2)), derivative_constraints=DB(2, 5),
                                    turning_constraint=TB(1.8, "angular_rate"), sfc_constraints=None,
                                    obstacle_constraints=[Ob(center=_col(5.5, 7), radius=1)])
    return 2, cc, dict()


def obstacles8(ns):
    """Config C2: the same shape with 8 circular obstacles."""
    W, WD, DB, TB, Ob = ns["Waypoint"], ns["WaypointData"], ns["DerivativeBounds"], ns["TurningBound"], ns["Obstacle"]
    w1 = W(location=_col(3, 4), velocity=_col(1, 0)); w2 = W(location=_col(9, 10), velocity=_col(1, 1))
    centers = [(5.5, 7), (4, 9), (8, 6), (6.5, 3.5), (2.5, 7.5), (7.5, 11.5), (10.5, 8), (5, 5.2)]
    radii = [1, 0.8, 0.9, 0.6, 0.7, 0.5, 0.8, 0.4]
    obs = [Ob(center=_col(*c), radius=r) for c, r in zip(centers, radii)]
    cc = ns["ConstraintsContainer"](waypoint_constraints=WD((w1, w2)), derivative_constraints=DB(2, 5),
                                    turning_constraint=TB(1.8, "angular_rate"), obstacle_constraints=obs)
    return 2, cc, dict()


def intermediate_waypoints(ns):
    """test_intermediate_waypoints.py:14-55 (config C3 shape, as shipped: angular rate bound 2)."""
    W, WD, DB, TB = ns["Waypoint"], ns["WaypointData"], ns["DerivativeBounds"], ns["TurningBound"]
    seq = (W(location=_col(3, 4), velocity=_col(1, 0)), W(location=_col(1.5, 6), velocity=_col(0, 1)),
           W(location=_col(3.4, 8), velocity=_col(4, 0)), W(location=_col(2, 10), velocity=_col(0, 1)))
    cc = ns["ConstraintsContainer"](waypoint_constraints=WD(seq), derivative_constraints=DB(5, None),
                                    turning_constraint=TB(2, "angular_rate"))
    return 2, cc, dict(num_intervals_free_space=14)


def intermediate_curvature(ns):
    """Config C3: intermediate waypoints with a curvature bound and a velocity bound."""
    W, WD, DB, TB = ns["Waypoint"], ns["WaypointData"], ns["DerivativeBounds"], ns["TurningBound"]
    seq = (W(location=_col(3, 4), velocity=_col(1, 0)), W(location=_col(5.5, 5), velocity=_col(1, 1)),
           W(location=_col(7, 7.5), velocity=_col(0, 1.5)), W(location=_col(6, 10), velocity=_col(-1, 1)))
    cc = ns["ConstraintsContainer"](waypoint_constraints=WD(seq), derivative_constraints=DB(5, None),
                                    turning_constraint=TB(2, "curvature"))
    return 2, cc, dict(num_intervals_free_space=14)


def sfc3d(ns):
    """test_sfc_trajectory_3D.py:19-89 (config C4 shape, shipped 3 corridors, ipc=[2,2,5])."""
    W, WD, DB = ns["Waypoint"], ns["WaypointData"], ns["DerivativeBounds"]
    pts = [_col(3, 4, 0), _col(7, 10, 3), _col(14, 7, 7), _col(20, 31, 20)]
    dims = [(3, 2, 3), (2, 3, 4), (3, 2, 2)]
    sfcs = []
    for i in range(3):
        R, T, L = ns["get3DRotationAndTranslationFromPoints"](pts[i], pts[i + 1])
        sfcs.append(ns["SFC"](np.array([[L + dims[i][0]], [dims[i][1]], [dims[i][2]]]), T, R))
    sfc = ns["SFC_Data"](tuple(sfcs), np.concatenate(pts, 1), 1)
    # positional DerivativeBounds(5, 0.3, None, None, None) as in the script (:60)
    wd = WD((W(location=pts[0], velocity=_col(1, 0, 0)), W(location=pts[3], velocity=_col(0, 0, 0))))
    cc = ns["ConstraintsContainer"](waypoint_constraints=wd, derivative_constraints=DB(5, 0.3, None, None, None),
                                    turning_constraint=None, sfc_constraints=sfc, obstacle_constraints=None)
    return 3, cc, dict(objective_function_type="minimal_velocity_path")


def sfc3d_four(ns):
    """Config C4: 4 corridors with near-equal segment lengths (ipc=[2,2,2,2], N=11)."""
    W, WD, DB = ns["Waypoint"], ns["WaypointData"], ns["DerivativeBounds"]
    pts = [_col(0, 0, 0), _col(7, 2, 1), _col(13, 7, 2), _col(20, 8, 5), _col(26, 13, 4)]
    dims = [(2.5, 2.2, 3), (2.2, 2.8, 3.5), (2.9, 2.1, 2.4), (2.4, 2.6, 3.1)]
    sfcs = []
    for i in range(4):
        R, T, L = ns["get3DRotationAndTranslationFromPoints"](pts[i], pts[i + 1])
        sfcs.append(ns["SFC"](np.array([[L + dims[i][0]], [dims[i][1]], [dims[i][2]]]), T, R))
    sfc = ns["SFC_Data"](tuple(sfcs), np.concatenate(pts, 1), 1)
    v0 = (pts[1] - pts[0]) / np.linalg.norm(pts[1] - pts[0])
    wd = WD((W(location=pts[0], velocity=v0), W(location=pts[4], velocity=_col(0, 0, 0))))
    cc = ns["ConstraintsContainer"](waypoint_constraints=wd, derivative_constraints=DB(5, 0.3),
                                    sfc_constraints=sfc)
    return 3, cc, dict(objective_function_type="minimal_velocity_path")


def sfc_obstacles3d(ns):
    """Coverage problem (not from a demo): the 4-corridor C4 shape with two spheres inside the corridors -- corridor
    rows AND rows behind them in one problem (the solver stores no corridor rows: the obstacle rows move up)."""
    Ob = ns["Obstacle"]
    d, cc, kw = sfc3d_four(ns)
    cc.obstacle_constraints = [Ob(center=_col(10, 4.5, 2.2), radius=0.5), Ob(center=_col(17, 8.5, 3), radius=0.6)]
    return d, cc, kw


def bicycle3(ns):
    """bicycle_trajectory_3.py (config C5, angular-rate variant)."""
    W, WD, DB, TB = ns["Waypoint"], ns["WaypointData"], ns["DerivativeBounds"], ns["TurningBound"]
    max_delta = 30 * np.pi / 180
    max_beta = np.arctan2(0.5 * np.tan(max_delta), 1)
    max_curv = np.tan(max_delta) * np.cos(max_beta) / 1
    wd = WD((W(location=_col(-5, 0), velocity=_col(0, 28)), W(location=_col(5, 0), velocity=_col(0, 20))))
    cc = ns["ConstraintsContainer"](wd, DB(30, 100), TB(max_curv * 30, "angular_rate"))
    return 2, cc, dict(objective_function_type="minimal_velocity_and_time_path", num_intervals_free_space=5)


def unicycle2(ns):
    """unicycle_trajectory_2.py shape (config C5, curvature variant)."""
    W, WD, DB, TB = ns["Waypoint"], ns["WaypointData"], ns["DerivativeBounds"], ns["TurningBound"]
    wd = WD((W(location=_col(-5, 0), velocity=_col(0, 28)), W(location=_col(5, 0), velocity=_col(0, 20))))
    cc = ns["ConstraintsContainer"](wd, DB(30, 100), TB(8 / 28, "curvature"))
    return 2, cc, dict(objective_function_type="minimal_velocity_and_time_path", num_intervals_free_space=5)


def bicycle_tangential(ns):
    """bicycle_trajectory.py:20-73 (tangential-acceleration rows + curvature)."""
    W, WD, DB, TB = ns["Waypoint"], ns["WaypointData"], ns["DerivativeBounds"], ns["TurningBound"]
    max_delta = 25 * np.pi / 180
    max_beta = np.arctan2(0.5 * np.tan(max_delta), 1)
    max_curv = np.tan(max_delta) * np.cos(max_beta) / 1
    wd = WD((W(location=_col(-5, 0), velocity=_col(0, 28)), W(location=_col(5, 0), velocity=_col(0, 28))))
    db = DB(30, max_tangential_acceleration=100, min_tangential_acceleration=-100)
    cc = ns["ConstraintsContainer"](wd, db, TB(max_curv, "curvature"))
    return 2, cc, dict(objective_function_type="minimal_velocity_and_time_path", num_intervals_free_space=8)


def features3d(ns):
    """Coverage problem (not from a demo): 3-D, start direction + acceleration, zero-velocity end with
    direction, every DerivativeBounds field, centripetal bound, two spheres, one intermediate waypoint."""
    W, WD, DB, TB, Ob = ns["Waypoint"], ns["WaypointData"], ns["DerivativeBounds"], ns["TurningBound"], ns["Obstacle"]
    w1 = W(location=_col(0, 0, 1), direction=_col(1, 0.2, 0), acceleration=_col(0.1, 0, 0))
    wm = W(location=_col(4, 3, 2))
    w2 = W(location=_col(9, 5, 3), velocity=_col(0, 0, 0), direction=_col(0, 1, 0))
    db = DB(max_velocity=6, max_acceleration=4, max_jerk=9, gravity=0.3, max_upward_velocity=3,
            max_horizontal_velocity=5.5, min_velocity=0.05)
    obs = [Ob(center=_col(2, 2.5, 1.2), radius=0.8), Ob(center=_col(7, 3, 3.5), radius=1.1)]
    cc = ns["ConstraintsContainer"](WD((w1, wm, w2)), db, TB(7, "centripetal_acceleration"), None, obs)
    return 3, cc, dict(objective_function_type="minimal_acceleration_and_time_path")


def features2d(ns):
    """Coverage problem: 2-D, zero-velocity start, target end waypoint, jerk + min-velocity bounds,
    up/horizontal bounds set in 2-D (rows stay zero), time-velocity-penalty objective."""
    W, WD, DB, TB = ns["Waypoint"], ns["WaypointData"], ns["DerivativeBounds"], ns["TurningBound"]
    w1 = W(location=_col(1, 1), velocity=_col(0, 0))
    w2 = W(location=_col(6, 4), velocity=_col(0.4, 0.1), is_target=True)
    db = DB(max_velocity=4, max_acceleration=3, max_jerk=8, max_upward_velocity=2, max_horizontal_velocity=3,
            min_velocity=0.01)
    cc = ns["ConstraintsContainer"](WD((w1, w2)), db, TB(3, "angular_rate"))
    return 2, cc, dict(objective_function_type="minimal_time_path_velocity_penalty")


def distance_time2d(ns):
    """Coverage problem: the one objective no demo uses, `minimal_distance_and_time_path` (TG/objectives/
    objective_functions.py:27-33), on the obstacle2d constraints."""
    d, cc, kw = obstacle2d(ns)
    return d, cc, dict(objective_function_type="minimal_distance_and_time_path")


def long2d(ns):
    """Coverage problem: more than 63 variables (2-D, num_intervals_free_space = 30 -> 33 control points, n = 67): the
    QP stage's lane-strided loops take two passes and its recurrences the shared-memory form."""
    d, cc, kw = obstacle2d(ns)
    return d, cc, dict(num_intervals_free_space=30)


ALL = dict(c1_sfc2d=c1_sfc2d, c1_curvature=c1_curvature, obstacle2d=obstacle2d, obstacles8=obstacles8,
           intermediate_waypoints=intermediate_waypoints, intermediate_curvature=intermediate_curvature,
           sfc3d=sfc3d, sfc3d_four=sfc3d_four, sfc_obstacles3d=sfc_obstacles3d, bicycle3=bicycle3, unicycle2=unicycle2,
           bicycle_tangential=bicycle_tangential, features3d=features3d, features2d=features2d,
           distance_time2d=distance_time2d, long2d=long2d)

# problems whose reference solve finishes in a few seconds (solve results are recorded for these)
SOLVE = ("c1_sfc2d", "c1_curvature", "obstacle2d", "obstacles8", "sfc3d", "sfc3d_four", "sfc_obstacles3d", "bicycle3",
         "unicycle2", "intermediate_waypoints", "distance_time2d", "long2d")


def test_point(x0, d, N, seed):
    """Evaluation point used for M1 parity: x0 + 0.3 N(0,1) on the control points,
    small positive perturbations on alpha / scalars, small shifts on the intermediate times."""
    rng = np.random.default_rng(seed)
    x = np.array(x0, dtype=float)
    x[:d * N] += 0.3 * rng.standard_normal(d * N)
    x[d * N] *= 1.0 + 0.2 * rng.random()
    for i in range(d * N + 1, len(x)):
        x[i] += 0.05 * rng.random() + 0.01
    return x
