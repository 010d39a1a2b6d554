"""Drop-in alias of the reference's module path (see trajectory_generator_b200/matrix_evaluation.py)."""
from trajectory_generator_b200.matrix_evaluation import *  # noqa: F401,F403
from trajectory_generator_b200.matrix_evaluation import (  # noqa: F401
    count_number_of_control_points, get_dimension, matrix_bspline_derivative_evaluation_for_dataset,
    matrix_bspline_derivative_evaluation_for_discrete_steps, matrix_bspline_evaluation_for_dataset,
    matrix_bspline_evaluation_for_discrete_steps, matrix_bspline_evaluation_for_timedataset)
