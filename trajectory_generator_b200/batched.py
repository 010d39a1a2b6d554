"""Batched front end: B problems of one shape from arrays, assembled, solved and sampled on the GPU
(per-container packing in Python costs 170-400 us, 30-80x the GPU solve).

``BatchedProblem`` takes what a ``ConstraintsContainer`` holds, as CUDA tensors with a leading batch axis, and produces
the same (shape descriptor, parameter rows, initial variables) ``problem.pack_problem`` produces per container:

    bp = BatchedProblem(dimension=2, start=s, end=g, start_velocity=v0, end_velocity=v1,
                        max_velocity=2.0, max_acceleration=5.0, turning=("angular_rate", 1.8),
                        obstacle_centers=ctr, obstacle_radii=rad)
    out = bp.solve()                      # dict(x, f, status, nit, violation); control points: bp.control_points(out)
    pos = bp.sample(out, num_points=200)  # [B, d, 200]

Supported blocks (the ones the BASELINE configurations use): terminal waypoints with velocities or a zero-velocity
end, intermediate waypoints with velocities, max velocity / max acceleration, one turning bound, spherical obstacles,
safe-flight corridors given by their end points and box paddings.  Everything else goes through the per-container API.
The parameter-row order is that of ``pack_problem`` (checked against it in tests/test_batched.py).
"""
import numpy as np

from . import batch as tgb, builder, matrix_evaluation, problem as pk


def _torch():
    import torch
    return torch


class BatchedProblem:
    def __init__(self, dimension, start, end, start_velocity=None, end_velocity=None, end_zero_velocity=False,
                 intermediate_locations=None, intermediate_velocities=None, max_velocity=None, max_acceleration=None,
                 turning=None, obstacle_centers=None, obstacle_radii=None, corridor_points=None, corridor_pads=None,
                 intervals_per_corridor=None, objective_function_type="minimal_velocity_and_time_path",
                 num_intervals_free_space=None, initial_scale_factor=1.0):
        torch = _torch()
        d = int(dimension)
        if not start.is_cuda:
            raise RuntimeError("BatchedProblem needs CUDA tensors (there is no CPU path)")
        if objective_function_type not in pk.OBJECTIVES:
            raise Exception("Error, Invalid objective function type")
        B, dev = start.shape[0], start.device
        col = lambda v: (torch.full((B, 1), float(v), dtype=torch.float64, device=dev) if np.isscalar(v)
                         else v.reshape(B, -1).to(torch.float64))
        niw = 0 if intermediate_locations is None else intermediate_locations.shape[2]
        ncorr = 0 if corridor_points is None else corridor_points.shape[2] - 1
        # ---- sizes (TG/trajectory_generator.py:134-162)
        if ncorr:
            ipc = [2] * ncorr if intervals_per_corridor is None else [int(v) for v in intervals_per_corridor]
            nint = sum(ipc)
        elif num_intervals_free_space is not None:
            nint = int(num_intervals_free_space)
        else:
            nint = 5 + 2 * int(bool(end_zero_velocity))
        N = nint + 3
        spec = np.zeros(pk.SP_COUNT, dtype=np.int32)
        spec[pk.SP_DIM], spec[pk.SP_NCP] = d, N
        spec[pk.SP_OBJECTIVE] = pk.OBJECTIVES.index(objective_function_type)
        par = [col(start), col(end)]
        if start_velocity is None:
            raise IndexError("terminal waypoint needs a velocity, direction or acceleration")
        spec[pk.SP_START_VEL] = 1
        par.append(col(start_velocity))
        if end_zero_velocity:
            spec[pk.SP_END_KIND] = 1
        else:
            if end_velocity is None:
                raise IndexError("terminal waypoint needs a velocity, direction or acceleration")
            spec[pk.SP_END_VEL] = 1
            par.append(col(end_velocity))
        if niw:
            spec[pk.SP_NIW] = niw
            par.append(col(intermediate_locations))                     # (d, niw) row-major per problem
            if intermediate_velocities is not None:
                spec[pk.SP_IW_VEL] = 1
                par.append(col(intermediate_velocities))
        if max_velocity is not None:
            spec[pk.SP_DB_MAXV] = 1
            par.append(col(max_velocity))
        if max_acceleration is not None:
            spec[pk.SP_DB_MAXA] = 1
            par.append(col(max_acceleration))
        if turning is not None:
            spec[pk.SP_TURN] = pk.TURN_KINDS[turning[0]]
            par.append(col(turning[1]))
        lay_sfc = None
        if ncorr:
            if ncorr > pk.MAX_CORRIDORS:
                raise Exception("at most %d corridors are supported" % pk.MAX_CORRIDORS)
            spec[pk.SP_NCORR] = ncorr
            spec[pk.SP_IPC0:pk.SP_IPC0 + ncorr] = ipc
            lay_sfc = len(par)
            par.append(torch.zeros((B, ncorr * (d * d + 2 * d)), dtype=torch.float64, device=dev))
        if obstacle_centers is not None:
            K = obstacle_centers.shape[1]
            spec[pk.SP_NOBST] = K
            par += [obstacle_centers.transpose(1, 2).reshape(B, -1).to(torch.float64), col(obstacle_radii)]   # c-major
        self.spec = spec
        self.layout = pk.Layout(spec)
        self.par = torch.cat(par, 1).contiguous()
        if self.par.shape[1] != self.layout.P:
            raise RuntimeError("parameter rows have %d entries, layout expects %d" % (self.par.shape[1], self.layout.P))
        if ncorr:
            builder.sfc_boxes_batch(spec, corridor_points.to(torch.float64), corridor_pads.to(torch.float64), self.par)
        # ---- initial variables (TG/objectives/objective_variables.py:27-48): along the corridor points, else the waypoints
        wseq = None
        if niw:
            wseq = torch.cat([start.reshape(B, d, 1), intermediate_locations, end.reshape(B, d, 1)], 2).to(torch.float64)
        seq = corridor_points.to(torch.float64) if ncorr else (
            wseq if wseq is not None else torch.stack([start, end], 2).to(torch.float64))
        self.x0 = builder.initial_guess_batch(spec, seq, wseq, initial_scale_factor)
        self.B, self.d, self.N = B, d, N

    def solve(self, jacobian="fd", maxiter=100, ftol=1e-6):
        """-> dict(x [B,n], f, status, nit, violation) of CUDA tensors (x is a fresh tensor; self.x0 is kept)."""
        x = self.x0.clone()
        return tgb.solve(self.spec, self.par, x, maxiter=maxiter, ftol=ftol, jacobian=jacobian)

    def control_points(self, out):
        """-> (control_points [B, d, N], scale_factors [B]) views of the solver rows"""
        x = out["x"]
        return x[:, :self.d * self.N].reshape(self.B, self.d, self.N), x[:, self.d * self.N]

    def sample(self, out, num_points=None, dt=None, derivative_order=0):
        return matrix_evaluation.sample_batch((out["x"], self.d, self.N), derivative_order=derivative_order,
                                              num_points=num_points, dt=dt)
