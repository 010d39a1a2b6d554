"""Alias of trajectory_generator_b200.path_generator under the reference's import path (TG/path_generator.py)."""
from trajectory_generator_b200.path_generator import PathGenerator  # noqa: F401
