"""Drop-in ``PathGenerator`` (reference TG/path_generator.py:31-225) on the CUDA path.

``generate_path`` keeps the reference signature and returns the optimised control points ``[d, N]``.  A path
problem is the trajectory problem with fewer rows -- plain location rows at both ends, direction rows, intermediate
locations, the curvature bound (as the turning row, or "indirect" as min velocity 0.5 / max acceleration
kappa 0.5^2), corridors and obstacles -- and one of the three time-free objectives; ``problem.pack_problem(...,
path_mode=...)`` writes it as a shape descriptor + parameter row for the same kernels that solve trajectory problems.
``generate_paths`` is the batched addition (containers grouped by shape, one library call).

Status: the mapping is checked on the CPU against fixtures recorded from the unmodified reference
(tests/test_path_generator.py: SLSQP-facing values to 1e-9 and converged control points to 1e-5 through the host
build of the kernel source); the kernels are the ones the trajectory problems already run on the GPU.
"""
import numpy as np

from . import batch
from .constraint_data_structures.constraints_container import ConstraintsContainer
from .constraint_data_structures.waypoint_data import Waypoint
from .problem import pack_problem


class PathGenerator:
    def __init__(self, dimension: int, jacobian: str = "fd", maxiter: int = 100, ftol: float = 1e-6):
        """jacobian / maxiter / ftol as for TrajectoryGenerator ("fd" follows the reference's iterates)."""
        self._dimension = dimension
        self._order = 3
        self._jacobian, self._maxiter, self._ftol = jacobian, maxiter, ftol
        self.last_result = None

    # ---- reference API ---------------------------------------------------------------------------
    def generate_path(self, constraints_container: ConstraintsContainer,
                      objective_function_type: str = "minimal_velocity_path",
                      num_intervals_free_space: int = None,
                      initial_control_points: np.ndarray = None,
                      initial_scale_factor: float = None,
                      isIndirect: bool = False):
        res = self.generate_paths([constraints_container], objective_function_type, num_intervals_free_space,
                                  [initial_control_points], [initial_scale_factor], isIndirect)[0]
        self.last_result = res
        return res["control_points"]

    def get_terminal_waypoint_properties(self, control_points: np.ndarray, scale_factor: float, side: str):
        """TG/path_generator.py:84-89 (the reference passes an unknown keyword to Waypoint there and raises
        TypeError; this returns what that method is meant to return)."""
        P = np.asarray(control_points, dtype=float)
        if side not in ("start", "end"):
            raise Exception("Funtion does not support this side value")
        a, b, c = (P[:, 0], P[:, 1], P[:, 2]) if side == "start" else (P[:, -3], P[:, -2], P[:, -1])
        return Waypoint(location=((a + 4 * b + c) / 6)[:, None], velocity=((c - a) / (2 * scale_factor))[:, None],
                        acceleration=((a - 2 * b + c) / (scale_factor * scale_factor))[:, None])

    # ---- batched addition ------------------------------------------------------------------------
    def generate_paths(self, containers, objective_function_type="minimal_velocity_path", num_intervals_free_space=None,
                       initial_control_points=None, initial_scale_factors=None, isIndirect=False):
        """Solves every container; returns one dict(control_points [d, N], status, nit, fun, x) per container, in
        input order.  Containers of different shapes are solved in one call (tg_solve_mixed_host)."""
        count = len(containers)
        icps = initial_control_points if initial_control_points is not None else [None] * count
        isfs = initial_scale_factors if initial_scale_factors is not None else [None] * count
        mode = "indirect" if isIndirect else "direct"
        packed = [pack_problem(self._dimension, cc, objective_function_type, num_intervals_free_space, icps[i], isfs[i],
                               path_mode=mode) for i, cc in enumerate(containers)]
        groups = {}
        for i, p in enumerate(packed):
            groups.setdefault(p.key, []).append(i)
        order = list(groups.values())
        buckets = [(packed[idx[0]].spec, np.stack([packed[i].par for i in idx]),
                    np.stack([np.clip(packed[i].x0, packed[i].xl, packed[i].xu) for i in idx])) for idx in order]
        if len(buckets) == 1:
            outs = [batch.solve_host(*buckets[0], self._maxiter, self._ftol, self._jacobian)]
        else:
            outs = batch.solve_mixed_host(buckets, self._maxiter, self._ftol, self._jacobian)
        results = [None] * count
        for idx, out in zip(order, outs):
            lay = packed[idx[0]].layout
            for k, i in enumerate(idx):
                x = out["x"][k]
                results[i] = dict(control_points=np.reshape(x[:lay.d * lay.N], (lay.d, lay.N)).copy(),
                                  status=int(out["status"][k]), nit=int(out["nit"][k]), fun=float(out["f"][k]), x=x.copy())
        return results
