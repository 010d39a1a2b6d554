"""B200-native batched B-spline trajectory optimisation (drop-in for trajectory_generator's hot path)."""
