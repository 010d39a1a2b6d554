from trajectory_generator_b200.path_plotter import set_axes_equal  # noqa: F401
