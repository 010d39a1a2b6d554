from trajectory_generator_b200.trajectory_generator import TrajectoryGenerator, TrajectoryResult  # noqa: F401
