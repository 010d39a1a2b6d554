"""Drop-in for the reference's ``SplineDataConcatenater`` (TG/spline_data_concatenater.py:5-33): samples a chain of
splines on one common clock with period dt; every spline is sampled on the GPU (matrix_evaluation.py)."""
import numpy as np

from .matrix_evaluation import (matrix_bspline_derivative_evaluation_for_discrete_steps,
                                matrix_bspline_evaluation_for_discrete_steps)


class SplineDataConcatenater:
    def __init__(self, dimension):
        self._dimension = dimension

    def concatenate_spline_data(self, dt, start_time, order_list, control_point_array_list, scale_factor_list,
                                derivative_order=0):
        pieces, clocks = [np.empty((self._dimension, 0))], [np.empty(0)]
        begin, offset = start_time, 0
        for order, cps, scale in zip(order_list, control_point_array_list, scale_factor_list):
            if derivative_order == 0:
                data, t, remainder, end = matrix_bspline_evaluation_for_discrete_steps(order, cps, begin, offset, dt, scale)
            else:
                data, t, remainder, end = matrix_bspline_derivative_evaluation_for_discrete_steps(
                    order, derivative_order, scale, cps, begin, offset, dt)
            begin, offset = end, dt - remainder          # the next spline starts where this one ends, on the same clock
            pieces.append(data); clocks.append(t)
        return np.concatenate(pieces, 1), np.concatenate(clocks)
