"""Drop-in for the reference's ``trajectory_generation/spline_order_converter.py``: converts a
B-spline into one of another order that follows it as closely as possible, with position, velocity and acceleration
matched at both ends.

    SmoothingSpline(order, dimension, resolution).generate_new_control_points(old_control_points, old_scale_factor,
                                                                              old_order, max_velocity=None)
        -> (optimized_control_points [d, N], scale_factor)            TG/spline_order_converter.py:22-34

Everything runs on the GPU: the old spline is sampled by the CUDA samplers (matrix_evaluation.sample_batch), the
arc-length initial guess (``create_initial_control_points``, :83-112) by ``tg_smooth_initial_batch`` and the SLSQP
solve by ``tg_smooth_batch`` (csrc/tg_smooth.h: the batched SQP stages of the trajectory solver with scipy's
finite-difference gradient).  ``generate_new_control_points_batch`` is the batched addition.  No CPU path.
"""
import ctypes

import numpy as np

from . import _native
from . import matrix_evaluation as me


class SmoothingSpline:
    """This class generates a new spline from a previous one"""

    def __init__(self, order, dimension, resolution):
        self._dimension = dimension
        self._resolution = resolution      # points spline
        self._order = order
        self.last_result = None

    # ---- reference API ---------------------------------------------------------------------------
    def generate_new_control_points(self, old_control_points, old_scale_factor, old_order, max_velocity=None):
        cps, scale = self.generate_new_control_points_batch(np.asarray(old_control_points, dtype=np.float64)[None],
                                                            [old_scale_factor], old_order)
        return cps[0], float(scale[0])

    def create_initial_control_points(self, old_pts, old_order, num_cont_pts):
        """TG/spline_order_converter.py:83-112 (equal arc-length steps along the old control polygon)."""
        torch = _torch()
        old = torch.from_numpy(np.ascontiguousarray(np.asarray(old_pts, dtype=np.float64))[None]).to(_device(torch))
        return self._initial(torch, old, int(num_cont_pts))[0].cpu().numpy()

    # ---- batched addition ------------------------------------------------------------------------
    def generate_new_control_points_batch(self, old_control_points, old_scale_factors, old_order):
        """old_control_points [B, d, oldN] (numpy or CUDA tensor, every spline with the same number of control points),
        old_scale_factors [B] -> (new control points [B, d, N], new scale factors [B]) as numpy arrays;
        ``last_result`` holds status / nit / fun per spline."""
        torch = _torch()
        dev = _device(torch)
        old = torch.as_tensor(np.asarray(old_control_points, dtype=np.float64) if not torch.is_tensor(old_control_points)
                              else old_control_points, dtype=torch.float64, device=dev).contiguous()
        B, d, oldN = old.shape
        if d != self._dimension:
            raise Exception("control points do not have dimension %d" % self._dimension)
        osf = torch.as_tensor(np.asarray(old_scale_factors, dtype=np.float64), dtype=torch.float64, device=dev).contiguous()
        k, R = int(self._order), int(self._resolution)
        old_int = oldN - int(old_order)
        N = int(old_int * 2.5) + k                                   # :71-75
        new_scale = old_int * osf / (N - k)                          # :77-82
        if not bool((new_scale == new_scale[0]).all()):
            raise NotImplementedError("one call converts splines of one scale factor (the new scale is a shape constant)")
        # objective data (:36-38) and end-point data (:47-50): samples of the old spline
        Y = me.sample_batch(old, num_points=R, order=old_order)                                     # [B, d, R]
        ends = [me.sample_batch(old, osf, derivative_order=r, num_points=2, order=old_order) for r in (0, 1, 2)]
        par = torch.cat([Y.reshape(B, d * R), torch.stack(ends, 2).reshape(B, d * 6)], 1).contiguous()   # b[c][r*2+e]
        x = self._initial(torch, old, N).reshape(B, d * N).contiguous()
        lib = _native.lib()
        with torch.cuda.device(dev):
            nbytes = lib.tg_smooth_workspace_bytes(d, N, k, R, B)
            if nbytes == 0:
                raise RuntimeError("spline order converter: unsupported shape (d * N = %d variables; at most 160; order 2..5)" % (d * N))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            f = torch.empty(B, dtype=torch.float64, device=dev)
            status = torch.empty(B, dtype=torch.int32, device=dev)
            nit = torch.empty(B, dtype=torch.int32, device=dev)
            p = lambda t: ctypes.c_void_p(t.data_ptr())
            rc = lib.tg_smooth_batch(d, N, k, R, float(new_scale[0].item()), B, p(par), p(x), p(f), p(status), p(nit), 100, 1e-6,
                                     p(ws), nbytes, ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _native.check(rc, "tg_smooth_batch")
        self.last_result = dict(status=status.cpu().numpy(), nit=nit.cpu().numpy(), fun=f.cpu().numpy())
        return x.reshape(B, d, N).cpu().numpy(), new_scale.cpu().numpy()

    def _initial(self, torch, old, N):
        B, d, oldN = old.shape
        x0 = torch.empty((B, d, N), dtype=torch.float64, device=old.device)
        scr = torch.empty((B, oldN), dtype=torch.float64, device=old.device)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(old.device):
            rc = _native.lib().tg_smooth_initial_batch(d, oldN, N, B, p(old), p(x0), p(scr),
                                                       ctypes.c_void_p(torch.cuda.current_stream(old.device).cuda_stream))
        _native.check(rc, "tg_smooth_initial_batch")
        return x0


def _torch():
    import torch
    return torch


def _device(torch):
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device available (this package has no CPU path)")
    return torch.device("cuda", torch.cuda.current_device())
