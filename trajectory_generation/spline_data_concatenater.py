"""Drop-in alias of the reference's module path (see trajectory_generator_b200/spline_data_concatenater.py)."""
from trajectory_generator_b200.spline_data_concatenater import SplineDataConcatenater  # noqa: F401
