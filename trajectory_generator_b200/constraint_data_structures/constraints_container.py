"""All constraints of one trajectory problem (reference DS/constraints_container.py:9-27)."""
from dataclasses import dataclass
from .dynamic_bounds import DerivativeBounds, TurningBound
from .obstacle import Obstacle, ObstacleList  # noqa: F401
from .safe_flight_corridor import SFC_Data
from .waypoint_data import WaypointData


@dataclass
class ConstraintsContainer:
    waypoint_constraints: WaypointData
    derivative_constraints: DerivativeBounds = None
    turning_constraint: TurningBound = None
    sfc_constraints: SFC_Data = None
    obstacle_constraints: 'list[Obstacle]' = None
