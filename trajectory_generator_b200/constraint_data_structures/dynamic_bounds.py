"""Kinematic bounds (reference DS/dynamic_bounds.py:5-65)."""
from dataclasses import dataclass

_TURNING_TYPES = ("angular_rate", "curvature", "centripetal_acceleration")


@dataclass
class TurningBound:
    max_turning_bound: float = None
    bound_type: str = None  # one of _TURNING_TYPES

    def __post_init__(self):
        if self.bound_type not in _TURNING_TYPES:
            raise Exception("Bound type must be either [angular_rate, curvature, centripetal_acceleration]")

    def checkIfTurningBoundActive(self):
        return self.max_turning_bound is not None

    def checkIfCurvatureBoundActive(self):
        return self.max_turning_bound is not None and self.bound_type == "curvature"


@dataclass
class DerivativeBounds:
    # NB the shipped demos construct this positionally, so field order matters
    max_velocity: float = None
    max_acceleration: float = None
    max_jerk: float = None
    gravity: float = None
    max_upward_velocity: float = None
    max_horizontal_velocity: float = None
    min_velocity: float = None
    min_tangential_acceleration: float = None
    max_tangential_acceleration: float = None

    def __post_init__(self):
        for value, what in ((self.max_upward_velocity, "upward"), (self.max_horizontal_velocity, "horizontal")):
            if value is None:
                continue
            if self.max_velocity is None:
                raise Exception("To set max %s velocity you need a general max velocity" % what)
            if self.max_velocity < value:
                raise Exception("Max %s velocity should be less than or equal to general max velocity" % what)

    def checkIfDerivativesActive(self):
        return any(v is not None for v in (self.max_velocity, self.max_acceleration, self.min_velocity, self.max_jerk))

    def checkIfTangentialAccelerationActive(self):
        return self.min_tangential_acceleration is not None and self.max_tangential_acceleration is not None
