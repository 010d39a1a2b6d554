from trajectory_generator_b200.constraint_data_structures.dynamic_bounds import *  # noqa: F401,F403
