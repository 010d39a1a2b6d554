from trajectory_generator_b200.constraint_data_structures.waypoint_data import *  # noqa: F401,F403
