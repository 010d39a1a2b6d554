"""Output sampling of solved trajectories on the CUDA path: drop-in for the reference's
``trajectory_generation/matrix_evaluation.py`` samplers (same names, arguments and return values) plus batched forms.

Reference functions mirrored (TG = trajectory_generation/ of the reference):
  matrix_bspline_evaluation_for_dataset               TG/matrix_evaluation.py:5-33
  matrix_bspline_evaluation_for_timedataset           TG/matrix_evaluation.py:35-64
  matrix_bspline_evaluation_for_discrete_steps        TG/matrix_evaluation.py:66-102
  matrix_bspline_derivative_evaluation_for_dataset    TG/matrix_evaluation.py:104-135
  matrix_bspline_derivative_evaluation_for_discrete_steps   TG/matrix_evaluation.py:137-173

Every sample is computed by ``tg_sample_batch`` (csrc/tg_sample.cu); there is no CPU path.  Only cubic splines
(order 3, the order ``TrajectoryGenerator`` produces, TG/trajectory_generator.py:48) are supported.
"""
import ctypes

import numpy as np

from . import _native


def _torch():
    import torch
    return torch


def _check_order(order):
    if order != 3:
        raise NotImplementedError("the CUDA samplers evaluate cubic B-splines (order 3) only, got order %r" % (order,))


def get_dimension(control_points):
    control_points = np.asarray(control_points)
    return 1 if control_points.ndim == 1 else control_points.shape[0]


def count_number_of_control_points(control_points):
    control_points = np.asarray(control_points)
    return len(control_points) if control_points.ndim == 1 else control_points.shape[1]


def sample_batch(control_points, scale_factors=None, derivative_order=0, num_points=None, dt=None, offsets=None,
                 out=None):
    """Samples B cubic trajectories at once on the GPU.

    control_points: [B, d, N] float64 CUDA tensor (or a [B, n] tensor of solver variable rows together with
    ``dims=(d, N)`` packed as ``(tensor, d, N)``); scale_factors: [B] tensor (required for derivatives and for
    discrete steps).  Exactly one of ``num_points`` (uniform in interval units, the *_for_dataset functions) or
    ``dt`` (the *_for_discrete_steps functions, with optional per-trajectory ``offsets``) must be given.

    Returns ``data`` [B, d, num_points] for num_points, or ``(data [B, d, cap], times [B, cap], counts [B])`` for dt,
    where only the first counts[b] samples of trajectory b are defined.
    """
    torch = _torch()
    if isinstance(control_points, tuple):
        rows, d, N = control_points
        cps, stride = rows, rows.shape[1]
        scale_t, scale_stride = rows[:, d * N:], rows.shape[1]
        B = rows.shape[0]
    else:
        cps = control_points.contiguous()
        B, d, N = cps.shape
        stride = d * N
        scale_t, scale_stride = (scale_factors.contiguous() if scale_factors is not None else None), 1
    if not cps.is_cuda:
        raise RuntimeError("sample_batch() needs CUDA tensors (there is no CPU path)")
    assert cps.dtype == torch.float64
    if (num_points is None) == (dt is None):
        raise ValueError("give exactly one of num_points and dt")
    if scale_t is None and (derivative_order > 0 or dt is not None):
        raise ValueError("scale_factors are required for derivatives and for discrete steps")
    dev = cps.device
    lib = _native.lib()
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    ptr = lambda t: ctypes.c_void_p(0 if t is None else t.data_ptr())
    with torch.cuda.device(dev):
        if num_points is not None:
            cap = int(num_points)
            data = out if out is not None else torch.empty((B, d, cap), dtype=torch.float64, device=dev)
            rc = lib.tg_sample_batch(d, N, B, ptr(cps), stride, ptr(scale_t), scale_stride, int(derivative_order), 0, cap,
                                     None, 0.0, ptr(data), cap, None, None, stream)
            _native.check(rc, "tg_sample_batch")
            return data
        sc = scale_t if scale_t.dim() == 1 else scale_t[:, 0]
        off = offsets.contiguous() if offsets is not None else None
        # capacity: the longest trajectory of the batch
        longest = float((sc * (N - 3)).max().item()) if off is None else float((sc * (N - 3) - off).max().item())
        cap = int(longest / dt) + 2
        data = torch.empty((B, d, cap), dtype=torch.float64, device=dev)
        times = torch.empty((B, cap), dtype=torch.float64, device=dev)
        counts = torch.zeros(B, dtype=torch.int32, device=dev)
        rc = lib.tg_sample_batch(d, N, B, ptr(cps), stride, ptr(scale_t), scale_stride, int(derivative_order), 1, 0,
                                 ptr(off), float(dt), ptr(data), cap, ptr(times), ptr(counts), stream)
        _native.check(rc, "tg_sample_batch")
        return data, times, counts


def _single(control_points):
    torch = _torch()
    cp = np.asarray(control_points, dtype=np.float64)
    if cp.ndim == 1:
        raise NotImplementedError("the CUDA samplers take d x N control points with d = 2 or 3")
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device available (this package has no CPU path)")
    return torch.from_numpy(np.ascontiguousarray(cp)[None]).cuda()


def matrix_bspline_evaluation_for_dataset(order, control_points, num_points):
    """TG/matrix_evaluation.py:5-33 -> spline_data[d, num_points]"""
    _check_order(order)
    return sample_batch(_single(control_points), num_points=int(num_points))[0].cpu().numpy()


def matrix_bspline_derivative_evaluation_for_dataset(order, derivative_order, scale_factor, control_points, num_points):
    """TG/matrix_evaluation.py:104-135 -> spline_derivative_data[d, num_points]"""
    _check_order(order)
    torch = _torch()
    cps = _single(control_points)
    sf = torch.tensor([float(scale_factor)], dtype=torch.float64, device=cps.device)
    if derivative_order > 3:
        return np.zeros((cps.shape[1], int(num_points)))
    return sample_batch(cps, sf, derivative_order=int(derivative_order), num_points=int(num_points))[0].cpu().numpy()


def _discrete(order, derivative_order, scale_factor, control_points, start_time, starting_offset, dt):
    _check_order(order)
    torch = _torch()
    cps = _single(control_points)
    sf = torch.tensor([float(scale_factor)], dtype=torch.float64, device=cps.device)
    off = torch.tensor([float(starting_offset)], dtype=torch.float64, device=cps.device)
    data, times, counts = sample_batch(cps, sf, derivative_order=int(derivative_order), dt=float(dt), offsets=off)
    ns = int(counts[0].item())
    num_intervals = cps.shape[2] - order
    duration = scale_factor * num_intervals
    time_data = times[0, :ns].cpu().numpy()
    last_time_sample = (ns - 1) * dt + starting_offset
    remainder_time = duration - last_time_sample
    return data[0, :, :ns].cpu().numpy(), time_data + start_time, remainder_time, duration + start_time


def matrix_bspline_evaluation_for_discrete_steps(order, control_points, start_time, starting_offset, dt, scale_factor):
    """TG/matrix_evaluation.py:66-102 -> (spline_data, time_data, remainder_time, spline_end_time)"""
    return _discrete(order, 0, scale_factor, control_points, start_time, starting_offset, dt)


def matrix_bspline_derivative_evaluation_for_discrete_steps(order, derivative_order, scale_factor, control_points,
                                                            start_time, starting_offset, dt):
    """TG/matrix_evaluation.py:137-173 -> (spline_derivative_data, time_data, remainder_time, spline_end_time)"""
    return _discrete(order, derivative_order, scale_factor, control_points, start_time, starting_offset, dt)


def matrix_bspline_evaluation_for_timedataset(order, control_points, time_data, scale_factor):
    """TG/matrix_evaluation.py:35-64.  The reference packs the samples that fall inside the spline at the front of
    the output and leaves zeros behind them; the same is done here (the in-range samples are evaluated on the GPU as
    one trajectory with its own time axis via discrete evaluation of each sample)."""
    _check_order(order)
    torch = _torch()
    cps = _single(control_points)
    d, N = cps.shape[1], cps.shape[2]
    t = np.asarray(time_data, dtype=np.float64) / scale_factor
    nint = N - order
    keep = (t >= 0) & (t <= nint)
    out = np.zeros((d, len(t)))
    tk = np.asarray(time_data, dtype=np.float64)[keep]
    if tk.size:
        # one pseudo-trajectory per sample: offset = the sample's time, a single step (dt larger than the spline)
        B = tk.size
        rows = cps.expand(B, d, N).contiguous()
        sf = torch.full((B,), float(scale_factor), dtype=torch.float64, device=cps.device)
        off = torch.from_numpy(tk).to(cps.device)
        data, _, _ = sample_batch(rows, sf, dt=float(scale_factor * nint + 1.0), offsets=off)
        out[:, :B] = data[:, :, 0].cpu().numpy().T
    return out
