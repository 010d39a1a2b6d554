"""Drop-in alias: the reference's import paths (``trajectory_generation.*``) bound to the B200 implementation in
``trajectory_generator_b200``, so that scripts written against the reference run unchanged
(e.g. ``from trajectory_generation.trajectory_generator import TrajectoryGenerator``)."""
