"""Problem-description dataclasses (drop-in for the reference's
``trajectory_generation.constraint_data_structures`` package).

Same class names, field order, defaults, helper-method names and validation
exceptions as the reference; matplotlib is only imported inside the plot
helpers so that the package works on head-less GPU boxes.
"""
from .waypoint_data import Waypoint, WaypointData, plot2D_waypoints, plot3D_waypoints
from .dynamic_bounds import DerivativeBounds, TurningBound
from .obstacle import Obstacle, ObstacleList, plot_2D_obstacle, plot_2D_obstacles, plot_3D_obstacle, plot_3D_obstacles
from .safe_flight_corridor import (SFC, SFC_Data, plot_sfc, plot_sfcs, plot_2D_sfc, plot_3D_sfc,
                                   get2DRotationAndTranslationFromPoints, get3DRotationAndTranslationFromPoints)
from .constraints_container import ConstraintsContainer
from .constraint_function_data import ConstraintFunctionData
