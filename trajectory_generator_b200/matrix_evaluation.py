"""Output sampling of solved trajectories on the CUDA path: drop-in for the reference's
``trajectory_generation/matrix_evaluation.py`` samplers (same names, arguments and return values) plus batched forms.

Reference functions mirrored (TG = trajectory_generation/ of the reference):
  matrix_bspline_evaluation_for_dataset               TG/matrix_evaluation.py:5-33
  matrix_bspline_evaluation_for_timedataset           TG/matrix_evaluation.py:35-64
  matrix_bspline_evaluation_for_discrete_steps        TG/matrix_evaluation.py:66-102
  matrix_bspline_derivative_evaluation_for_dataset    TG/matrix_evaluation.py:104-135
  matrix_bspline_derivative_evaluation_for_discrete_steps   TG/matrix_evaluation.py:137-173

  evaluate_point_on_interval / evaluate_point_derivative_on_interval   TG/matrix_evaluation.py:183-195
  get_M_matrix, get_T_vector, get_T_derivative_vector                  TG/matrix_evaluation.py:197-262 (small constants /
                                                                       coefficient vectors, returned as numpy arrays)

Every sample is computed by ``tg_sample_batch_order`` / ``tg_interval_points_batch`` (csrc/tg_sample.cu); there is no
CPU path.  B-splines of order 2 .. 5 are supported (``TrajectoryGenerator`` produces order 3); like the reference,
order 1 raises (its ``get_M_matrix`` falls through to the ``else`` branch, TG/matrix_evaluation.py:203-214).
"""
import ctypes

import numpy as np

from . import _native


def _torch():
    import torch
    return torch


def _check_order(order):
    if order > 5:
        raise Exception("Error: Cannot compute higher than 5th order matrix evaluation")
    if order not in (2, 3, 4, 5):
        raise Exception("Cannot return M matrix for spline of order ", order)


def get_M_matrix(order):
    """TG/matrix_evaluation.py:197-262 (a constant per order)."""
    if order == 0:
        return 1
    _check_order(order)
    if order == 2:
        return .5 * np.array([[1, -2, 1], [-2, 2, 1], [1, 0, 0]])
    if order == 3:
        return np.array([[-2, 6, -6, 2], [6, -12, 0, 8], [-6, 6, 6, 2], [2, 0, 0, 0]]) / 12
    if order == 4:
        return np.array([[1, -4, 6, -4, 1], [-4, 12, -6, -12, 11], [6, -12, -6, 12, 11], [-4, 4, 6, 4, 1],
                         [1, 0, 0, 0, 0]]) / 24
    return np.array([[-1, 5, -10, 10, -5, 1], [5, -20, 20, 20, -50, 26], [-10, 30, 0, -60, 0, 66],
                     [10, -20, -20, 20, 50, 26], [-5, 5, 10, 10, 5, 1], [1, 0, 0, 0, 0, 0]]) / 120


def get_T_vector(order, t, tj, scale_factor):
    """TG/matrix_evaluation.py:224-232 -> [order+1, 1] powers of (t - tj) / scale_factor."""
    T = np.ones((order + 1, 1))
    t_tj = t - tj
    for i in range(order + 1):
        T[i, 0] = (t_tj / scale_factor) ** (order - i)
    return T


def get_T_derivative_vector(order, t, tj, rth_derivative, scale_factor):
    """TG/matrix_evaluation.py:216-222."""
    from math import factorial
    T = np.zeros((order + 1, 1))
    t_tj = t - tj
    for i in range(order - rth_derivative + 1):
        T[i, 0] = (t_tj ** (order - rth_derivative - i)) / (scale_factor ** (order - i)) * factorial(order - i) / factorial(order - i - rth_derivative)
    return T


def interval_points_batch(control_points, t, tj, scale_factors, derivative_order=0):
    """control_points [B, d, order+1], t / tj / scale_factors [B] (float64 CUDA tensors) -> [B, d]: one point (or
    r-th derivative) of one interval per item (tg_interval_points_batch)."""
    torch = _torch()
    cps = control_points.contiguous()
    if not cps.is_cuda:
        raise RuntimeError("interval_points_batch() needs CUDA tensors (there is no CPU path)")
    B, d, k = cps.shape
    out = torch.empty((B, d), dtype=torch.float64, device=cps.device)
    ptr = lambda x: ctypes.c_void_p(x.contiguous().data_ptr())
    with torch.cuda.device(cps.device):
        rc = _native.lib().tg_interval_points_batch(k - 1, d, B, ptr(cps), ptr(t), ptr(tj), ptr(scale_factors),
                                                    int(derivative_order), ptr(out),
                                                    ctypes.c_void_p(torch.cuda.current_stream(cps.device).cuda_stream))
    _native.check(rc, "tg_interval_points_batch")
    return out


def _one_point(control_points, t, tj, scale_factor, r):
    torch = _torch()
    cp = np.atleast_2d(np.asarray(control_points, dtype=np.float64))
    _check_order(cp.shape[1] - 1)
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device available (this package has no CPU path)")
    dev = torch.device("cuda", torch.cuda.current_device())
    f = lambda v: torch.tensor([float(v)], dtype=torch.float64, device=dev)
    out = interval_points_batch(torch.from_numpy(np.ascontiguousarray(cp)[None]).to(dev), f(t), f(tj), f(scale_factor), r)
    return out[0].cpu().numpy()[:, None]


def evaluate_point_on_interval(control_points, t, tj, scale_factor):
    """TG/matrix_evaluation.py:183-188 -> point [d, 1]"""
    return _one_point(control_points, t, tj, scale_factor, 0)


def evaluate_point_derivative_on_interval(control_points, t, tj, scale_factor, rth_derivative):
    """TG/matrix_evaluation.py:190-195 -> point [d, 1]"""
    return _one_point(control_points, t, tj, scale_factor, int(rth_derivative))


def get_dimension(control_points):
    control_points = np.asarray(control_points)
    return 1 if control_points.ndim == 1 else control_points.shape[0]


def count_number_of_control_points(control_points):
    control_points = np.asarray(control_points)
    return len(control_points) if control_points.ndim == 1 else control_points.shape[1]


def sample_batch(control_points, scale_factors=None, derivative_order=0, num_points=None, dt=None, offsets=None,
                 out=None, order=3):
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
            rc = lib.tg_sample_batch_order(int(order), d, N, B, ptr(cps), stride, ptr(scale_t), scale_stride, int(derivative_order), 0, cap,
                                     None, 0.0, ptr(data), cap, None, None, stream)
            _native.check(rc, "tg_sample_batch_order")
            return data
        sc = scale_t if scale_t.dim() == 1 else scale_t[:, 0]
        off = offsets.contiguous() if offsets is not None else None
        # capacity: the longest trajectory of the batch
        longest = float((sc * (N - order)).max().item()) if off is None else float((sc * (N - order) - off).max().item())
        cap = int(longest / dt) + 2
        data = torch.empty((B, d, cap), dtype=torch.float64, device=dev)
        times = torch.empty((B, cap), dtype=torch.float64, device=dev)
        counts = torch.zeros(B, dtype=torch.int32, device=dev)
        rc = lib.tg_sample_batch_order(int(order), d, N, B, ptr(cps), stride, ptr(scale_t), scale_stride, int(derivative_order), 1, 0,
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
    return sample_batch(_single(control_points), num_points=int(num_points), order=order)[0].cpu().numpy()


def matrix_bspline_derivative_evaluation_for_dataset(order, derivative_order, scale_factor, control_points, num_points):
    """TG/matrix_evaluation.py:104-135 -> spline_derivative_data[d, num_points]"""
    _check_order(order)
    torch = _torch()
    cps = _single(control_points)
    sf = torch.tensor([float(scale_factor)], dtype=torch.float64, device=cps.device)
    if derivative_order > order:          # K has no non-zero entry: the reference returns zeros
        return np.zeros((cps.shape[1], int(num_points)))
    if derivative_order > 3:
        raise NotImplementedError("derivative orders above 3 are not provided by the CUDA samplers")
    return sample_batch(cps, sf, derivative_order=int(derivative_order), num_points=int(num_points), order=order)[0].cpu().numpy()


def _discrete(order, derivative_order, scale_factor, control_points, start_time, starting_offset, dt):
    _check_order(order)
    torch = _torch()
    cps = _single(control_points)
    sf = torch.tensor([float(scale_factor)], dtype=torch.float64, device=cps.device)
    off = torch.tensor([float(starting_offset)], dtype=torch.float64, device=cps.device)
    data, times, counts = sample_batch(cps, sf, derivative_order=int(derivative_order), dt=float(dt), offsets=off, order=order)
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
        data, _, _ = sample_batch(rows, sf, dt=float(scale_factor * nint + 1.0), offsets=off, order=order)
        out[:, :B] = data[:, :, 0].cpu().numpy().T
    return out
