"""Waypoints (reference DS/waypoint_data.py:5-156)."""
from dataclasses import dataclass
import numpy as np


def _flat_len(a):
    return len(np.asarray(a).flatten())


@dataclass
class Waypoint:
    # field order is part of the public API (positional construction)
    location: np.ndarray
    direction: np.ndarray = None
    velocity: np.ndarray = None
    acceleration: np.ndarray = None
    jerk: np.ndarray = None
    dimension: int = None
    side: str = None
    is_target: bool = None

    def __post_init__(self):
        self.dimension = _flat_len(self.location)
        for name, label in (("velocity", "Velocity"), ("acceleration", "Acceleration"), ("jerk", "Jerk")):
            value = getattr(self, name)
            if value is not None and _flat_len(value) != self.dimension:
                raise Exception("Error: %s is not the same dimension as location" % label)
        # a non-zero velocity wins over a direction (reference DS/waypoint_data.py:47-51)
        if self.direction is not None and self.velocity is not None and np.linalg.norm(self.velocity) > 0:
            self.direction = None
            print("Using velocity constraint - cannot use both velocity and direction constraint")

    def checkIfDirectionActive(self):
        return self.direction is not None

    def checkIfVelocityActive(self):
        return self.velocity is not None

    def checkIfAccelerationActive(self):
        return self.acceleration is not None

    def checkIfZeroVel(self):
        return bool(self.velocity is not None and np.linalg.norm(self.velocity) <= 0)

    def checkIfDerivativesActive(self):
        # The reference tests the bound method ``self.checkIfVelocityActive``
        # (always truthy, DS/waypoint_data.py:19), so every waypoint that is
        # not a zero-velocity waypoint counts as having active derivatives.
        # Kept: it decides which constraint blocks exist.
        if self.checkIfAccelerationActive() or self.checkIfDirectionActive():
            return True
        return not self.checkIfZeroVel()


class WaypointData:
    """Start / end waypoint plus stacked intermediate locations and velocities."""

    def __init__(self, waypoint_sequence: 'list[Waypoint]'):
        if len(waypoint_sequence) < 2:
            raise Exception("Waypoint sequence must have at least two waypoints")
        first, last = waypoint_sequence[0], waypoint_sequence[-1]
        if first.dimension != last.dimension:
            raise Exception("Waypoint dimensions do not match")
        self.start_waypoint, self.end_waypoint = first, last
        self.dimension = first.dimension
        first.side, last.side = "start", "end"
        self.intermediate_locations = None
        self.intermediate_velocities = None
        middle = list(waypoint_sequence[1:-1])
        if middle:
            locs = np.zeros((self.dimension, len(middle)))
            vels = np.zeros((self.dimension, len(middle)))
            any_velocity = False
            for i, wp in enumerate(middle):
                if wp.dimension != self.dimension:
                    raise Exception("Waypoint dimensions do not match")
                locs[:, i] = wp.location.flatten()
                if wp.velocity is not None:
                    any_velocity = True
                    vels[:, i] = wp.velocity.flatten()
            self.intermediate_locations = locs
            self.intermediate_velocities = vels if any_velocity else None

    def get_waypoint_locations(self):
        parts = [self.start_waypoint.location]
        if self.intermediate_locations is not None:
            parts.append(self.intermediate_locations)
        parts.append(self.end_waypoint.location)
        return np.concatenate(parts, 1)

    def get_num_intermediate_waypoints(self):
        return 0 if self.intermediate_locations is None else np.shape(self.intermediate_locations)[1]

    def get_num_waypoint_scalars(self):
        return int(self.start_waypoint.direction is not None) + int(self.end_waypoint.direction is not None)


def _quiver_args(wp):
    return [wp.location.item(i) for i in range(wp.dimension)] + [wp.velocity.item(i) for i in range(wp.dimension)]


def plot2D_waypoints(waypoint_data: WaypointData, ax):
    pts = waypoint_data.get_waypoint_locations()
    ax.scatter(pts[0, :], pts[1, :], facecolors='none', edgecolors="r", label="waypoint constraints")
    for wp in (waypoint_data.start_waypoint, waypoint_data.end_waypoint):
        if wp.checkIfVelocityActive() and not wp.checkIfZeroVel():
            ax.quiver(*_quiver_args(wp), color="r")
    if waypoint_data.intermediate_locations is not None:
        mid = waypoint_data.intermediate_locations
        ax.scatter(mid[0, :], mid[1, :], facecolors='none', edgecolors="r")


def plot3D_waypoints(waypoint_data: WaypointData, ax):
    pts = waypoint_data.get_waypoint_locations()
    ax.scatter(pts[0, :], pts[1, :], pts[2, :], color="b")
    span = get_distance_between_start_and_end_waypoint(waypoint_data.start_waypoint, waypoint_data.end_waypoint)
    for wp in (waypoint_data.start_waypoint, waypoint_data.end_waypoint):
        if wp.checkIfVelocityActive():
            ax.quiver(*_quiver_args(wp), length=span / 10, normalize=True)
    if waypoint_data.intermediate_locations is not None:
        mid = waypoint_data.intermediate_locations
        ax.scatter(mid[0, :], mid[1, :], mid[2, :], color="b")


def get_distance_between_start_and_end_waypoint(start_waypoint, end_waypoint):
    return np.linalg.norm(end_waypoint.location - start_waypoint.location)
