from trajectory_generator_b200.constraint_data_structures.safe_flight_corridor import *  # noqa: F401,F403
