"""Batched problem construction on the GPU.

``initial_guess_batch``  -- the reference's create_initial_objective_variables (TG/objectives/objective_variables.py:
27-48, 63-105) for B problems of one shape: control points along the point sequence, scale factor, waypoint scalars,
intermediate-waypoint times.
``sfc_boxes_batch``      -- get2D/3DRotationAndTranslationFromPoints + SFC.getRotatedBounds
(DS/safe_flight_corridor.py:109-146, 13-16) for every corridor of every problem, written straight into the parameter
rows the solver reads.

CUDA tensors in, CUDA tensors out; there is no CPU path.
"""
import ctypes

from . import _native
from .problem import Layout


def _torch():
    import torch
    return torch


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def initial_guess_batch(spec, point_sequences, waypoint_sequences=None, initial_scale_factor=1.0, out=None):
    """point_sequences [B, d, npts] (waypoint locations, or the corridor point sequence when there are corridors);
    waypoint_sequences [B, d, niw + 2] when the shape has intermediate waypoints.  Returns x0 [B, n]."""
    torch = _torch()
    lay = Layout(spec)
    seq = point_sequences.contiguous()
    if not seq.is_cuda:
        raise RuntimeError("initial_guess_batch() needs CUDA tensors (there is no CPU path)")
    B, d, npts = seq.shape
    assert d == lay.d and seq.dtype == torch.float64
    wseq = waypoint_sequences.contiguous() if waypoint_sequences is not None else None
    nwp = wseq.shape[2] if wseq is not None else 0
    x0 = out if out is not None else torch.empty((B, lay.n), dtype=torch.float64, device=seq.device)
    spec, sp = _native.spec_ptr(spec)
    with torch.cuda.device(seq.device):
        rc = _native.lib().tg_initial_guess_batch(sp, B, _ptr(seq), npts, _ptr(wseq), nwp, float(initial_scale_factor),
                                                  _ptr(x0), ctypes.c_void_p(torch.cuda.current_stream(seq.device).cuda_stream))
    _native.check(rc, "tg_initial_guess_batch")
    return x0


def sfc_boxes_batch(spec, points, pads, par):
    """points [B, d, ncorr + 1]: consecutive corridor end points; pads [B, ncorr, d]: box dimensions beyond the segment
    length (dimension 0) / total (other dimensions); par [B, P]: parameter rows, whose corridor slots are filled in
    place.  Returns the segment lengths [B, ncorr]."""
    torch = _torch()
    lay = Layout(spec)
    points = points.contiguous(); pads = pads.contiguous()
    if not (points.is_cuda and par.is_cuda and pads.is_cuda):
        raise RuntimeError("sfc_boxes_batch() needs CUDA tensors (there is no CPU path)")
    B, d, np1 = points.shape
    assert d == lay.d and par.is_contiguous() and par.shape == (B, lay.P) and pads.shape == (B, np1 - 1, d)
    lengths = torch.empty((B, np1 - 1), dtype=torch.float64, device=points.device)
    spec, sp = _native.spec_ptr(spec)
    with torch.cuda.device(points.device):
        rc = _native.lib().tg_sfc_boxes_batch(sp, B, _ptr(points), _ptr(pads), _ptr(par), _ptr(lengths),
                                              ctypes.c_void_p(torch.cuda.current_stream(points.device).cuda_stream))
    _native.check(rc, "tg_sfc_boxes_batch")
    return lengths


def sfc_intervals_batch(points, min_intervals_per_corridor=1):
    """points [B, d, ncorr + 1] (CUDA): intervals per corridor as the reference's SFC_Data chooses them from the geometry
    (DS/safe_flight_corridor.py:78-88).  Returns (ipc [B, ncorr] int32, key [B] int64: equal for problems of one shape)."""
    torch = _torch()
    points = points.contiguous()
    if not points.is_cuda:
        raise RuntimeError("sfc_intervals_batch() needs CUDA tensors (there is no CPU path)")
    B, d, np1 = points.shape
    ipc = torch.empty((B, np1 - 1), dtype=torch.int32, device=points.device)
    key = torch.empty(B, dtype=torch.int64, device=points.device)
    with torch.cuda.device(points.device):
        rc = _native.lib().tg_sfc_intervals_batch(d, np1 - 1, B, _ptr(points), int(min_intervals_per_corridor), _ptr(ipc), _ptr(key),
                                                  ctypes.c_void_p(torch.cuda.current_stream(points.device).cuda_stream))
    _native.check(rc, "tg_sfc_intervals_batch")
    return ipc, key
