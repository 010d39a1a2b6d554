"""Problem definitions for PathGenerator (reference TG/path_generator.py:31-225), shared by the fixture generator
(reference classes) and the tests (this repo's classes): ``(dimension, container, generate_path kwargs)``."""
import numpy as np

from problems import _col, sfc3d_four


def directions_curvature(ns):
    W, WD, TB = ns["Waypoint"], ns["WaypointData"], ns["TurningBound"]
    wd = WD((W(location=_col(0, 0), direction=_col(1, 0)), W(location=_col(8, 3), direction=_col(0, 1))))
    return 2, ns["ConstraintsContainer"](waypoint_constraints=wd, turning_constraint=TB(0.6, "curvature")), dict()


def velocities_ignored_obstacle(ns):
    """terminal velocities are not constrained by generate_path (only locations and directions are)"""
    W, WD, TB, Ob = ns["Waypoint"], ns["WaypointData"], ns["TurningBound"], ns["Obstacle"]
    wd = WD((W(location=_col(0, 0), velocity=_col(1, 0)), W(location=_col(8, 3), velocity=_col(0, 1))))
    cc = ns["ConstraintsContainer"](waypoint_constraints=wd, turning_constraint=TB(1.0, "curvature"),
                                    obstacle_constraints=[Ob(center=_col(4, 1.4), radius=0.7)])
    return 2, cc, dict(objective_function_type="minimal_distance_path")


def intermediate_jerk(ns):
    W, WD = ns["Waypoint"], ns["WaypointData"]
    wd = WD((W(location=_col(0, 0), velocity=_col(1, 0)), W(location=_col(4, 4)), W(location=_col(8, 3), velocity=_col(0, 1))))
    return 2, ns["ConstraintsContainer"](waypoint_constraints=wd), dict(objective_function_type="minimal_acceleration_path")


def indirect_curvature(ns):
    """isIndirect: min velocity 0.5 and max acceleration kappa 0.5^2 instead of the curvature row"""
    W, WD, TB = ns["Waypoint"], ns["WaypointData"], ns["TurningBound"]
    wd = WD((W(location=_col(0, 0), velocity=_col(1, 0)), W(location=_col(6, 5), velocity=_col(0, 1))))
    return 2, ns["ConstraintsContainer"](waypoint_constraints=wd, turning_constraint=TB(0.8, "curvature")), dict(isIndirect=True)


def corridors3d(ns):
    d, cc, kw = sfc3d_four(ns)
    cc.derivative_constraints = None        # generate_path never reads them
    return d, cc, dict(objective_function_type="minimal_velocity_path")


def direction3d_zero_velocity(ns):
    """3-D, zero-velocity start (plain location row in a path problem, two more intervals) with a direction"""
    W, WD = ns["Waypoint"], ns["WaypointData"]
    wd = WD((W(location=_col(0, 0, 0), velocity=_col(0, 0, 0), direction=_col(1, 0, 0.2)),
             W(location=_col(6, 2, 3), direction=_col(0, 1, 0))))
    return 3, ns["ConstraintsContainer"](waypoint_constraints=wd), dict(objective_function_type="minimal_velocity_path")


ALL = dict(directions_curvature=directions_curvature, velocities_ignored_obstacle=velocities_ignored_obstacle,
           intermediate_jerk=intermediate_jerk, indirect_curvature=indirect_curvature, corridors3d=corridors3d,
           direction3d_zero_velocity=direction3d_zero_velocity)
