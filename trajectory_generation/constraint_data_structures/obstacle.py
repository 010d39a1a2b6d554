from trajectory_generator_b200.constraint_data_structures.obstacle import *  # noqa: F401,F403
