"""B200-native batched B-spline trajectory optimisation (drop-in for trajectory_generator's hot path)."""
from .constraint_data_structures import (ConstraintsContainer, DerivativeBounds, Obstacle, SFC, SFC_Data, TurningBound,  # noqa: F401
                                         Waypoint, WaypointData, get2DRotationAndTranslationFromPoints,
                                         get3DRotationAndTranslationFromPoints)
from .path_generator import PathGenerator  # noqa: F401
from .trajectory_generator import TrajectoryGenerator  # noqa: F401
