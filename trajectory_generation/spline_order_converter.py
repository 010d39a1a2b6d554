"""Drop-in alias of the reference's module path (see trajectory_generator_b200/spline_order_converter.py)."""
from trajectory_generator_b200.spline_order_converter import SmoothingSpline  # noqa: F401
