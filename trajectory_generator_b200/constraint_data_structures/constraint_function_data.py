"""Post-solve violation bookkeeping (reference DS/constraint_function_data.py:5-79)."""
from dataclasses import dataclass
import numpy as np
import numpy.typing as npt

_CLASSES = frozenset((
    "Derivative", "Obstacle", "Safe_Flight_Corridor", "Turning",
    "Start_Waypoint_Location", "End_Waypoint_Location",
    "Start_Waypoint_Derivatives", "End_Waypoint_Derivatives",
    "Start_Waypoint_Direction", "End_Waypoint_Direction",
    "Intermediate_Waypoint_Locations", "Intermediate_Waypoint_Velocities",
    "Zero_Velocity_End_Waypoint_Location", "Zero_Velocity_Start_Waypoint_Location",
    "Target_Location", "Target_Orbit_Location"))


@dataclass
class ConstraintFunctionData:
    constraint_function: callable
    lower_bound: npt.NDArray[np.float64]
    upper_bound: npt.NDArray[np.float64]
    key: npt.NDArray[np.dtype('U1')] = None
    constraint_class: str = None
    constraint_tolerance: float = 10e-6

    def __post_init__(self):
        if self.constraint_class not in _CLASSES:
            raise Exception("Constraint class [", self.constraint_class, "] invalid")

    def get_output(self, optimized_result):
        return self.constraint_function(optimized_result)

    def get_violations(self, output):
        tol = self.constraint_tolerance
        return np.logical_or(output > (self.upper_bound + tol), output < (self.lower_bound - tol))

    def get_error(self, output):
        cls = self.constraint_class
        if cls in ("Derivative", "Turning"):
            return output
        if cls == "Obstacle":
            return -output
        if cls == "Safe_Flight_Corridor":
            return np.max((self.lower_bound - output, output - self.upper_bound), 0)
        # every remaining class is an equality block: distance to its target value
        return np.abs(output - self.lower_bound)
