"""Safe flight corridors: oriented boxes (reference DS/safe_flight_corridor.py:5-146)."""
from dataclasses import dataclass
import numpy as np


@dataclass
class SFC:
    """Box of size ``dimensions`` centred at ``rotation @ translation``, axes = columns of ``rotation``."""
    dimensions: np.ndarray
    translation: np.ndarray
    rotation: np.ndarray

    def getRotatedBounds(self):
        half = self.dimensions / 2
        return self.translation - half, self.translation + half

    def getPointsToPlot(self):
        return self.getPointsToPlot2D() if len(self.dimensions.flatten()) == 2 else self.getPointsToPlot3D()

    def getPointsToPlot2D(self):
        lo, hi = self.getRotatedBounds()
        x0, x1, y0, y1 = lo.item(0), hi.item(0), lo.item(1), hi.item(1)
        outline = np.array([[x0, x0, x1, x1, x0],
                            [y0, y1, y1, y0, y0]])
        return self.rotation @ outline

    def getPointsToPlot3D(self):
        lo, hi = self.getRotatedBounds()
        x = (lo.item(0), hi.item(0)); y = (lo.item(1), hi.item(1)); z = (lo.item(2), hi.item(2))
        # one continuous pen stroke over all 12 edges (same vertex walk as the reference)
        walk = [(1, 0, 0), (0, 0, 0), (0, 1, 0), (1, 1, 0), (1, 1, 1), (0, 1, 1), (0, 0, 1), (1, 0, 1),
                (1, 0, 0), (1, 1, 0), (1, 1, 1), (1, 0, 1), (0, 0, 1), (0, 0, 0), (0, 1, 0), (0, 1, 1)]
        pts = np.array([[x[i], y[j], z[k]] for i, j, k in walk]).T
        return self.rotation @ pts


class SFC_Data:
    def __init__(self, sfc_list: list, point_sequence: np.ndarray, min_num_intervals_per_corridor: int = 1,
                 intervals_per_corridor=None):
        self._sfc_list = sfc_list
        self._num_corridors = len(sfc_list)
        self._point_sequence = point_sequence
        self._min_num_intervals_per_corridor = min_num_intervals_per_corridor
        if intervals_per_corridor is None:
            intervals_per_corridor = self.__evaluate_intervals_per_corridor()
        self._intervals_per_corridor = intervals_per_corridor
        self._num_intervals = np.sum(self._intervals_per_corridor)

    def get_sfc_list(self):
        return self._sfc_list

    def get_num_corridors(self):
        return self._num_corridors

    def get_point_sequence(self):
        return self._point_sequence

    def get_intervals_per_corridor(self):
        return self._intervals_per_corridor

    def get_num_intervals(self):
        return self._num_intervals

    def __evaluate_intervals_per_corridor(self):
        # reference DS/safe_flight_corridor.py:78-88: one corridor -> 5 intervals,
        # otherwise proportional to segment length relative to the shortest one
        if self._num_corridors < 2:
            return 5
        seg = np.linalg.norm(self._point_sequence[:, 1:] - self._point_sequence[:, :-1], 2, 0)
        shortest = np.min(seg)
        return [(int(np.round(seg[i] / shortest)) + 1) * self._min_num_intervals_per_corridor
                for i in range(self._num_corridors)]


def plot_sfc(sfc: SFC, ax):
    (plot_2D_sfc if len(sfc.dimensions.flatten()) == 2 else plot_3D_sfc)(sfc, ax)


def plot_sfcs(sfcs: list, ax):
    if sfcs is not None:
        for sfc in sfcs:
            plot_sfc(sfc, ax)


def plot_2D_sfc(sfc: SFC, ax):
    pts = sfc.getPointsToPlot()
    ax.plot(pts[0, :], pts[1, :])


def plot_3D_sfc(sfc: SFC, ax):
    pts = sfc.getPointsToPlot()
    ax.plot(pts[0, :], pts[1, :], pts[2, :])


def _rot_z(psi, dim):
    c, s = np.cos(psi), np.sin(psi)
    if dim == 2:
        return np.array([[c, -s], [s, c]])
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def get2DRotationAndTranslationFromPoints(point_1, point_2):
    """Rotation taking +x onto the segment direction, box centre in the box frame, segment length."""
    delta = point_2 - point_1
    rotation = _rot_z(np.arctan2(delta.item(1), delta.item(0)), 2)
    translation = rotation.T @ (point_1 + point_2) / 2
    return rotation, translation, np.linalg.norm(delta, 2)


def get3DRotationAndTranslationFromPoints(point_1, point_2):
    delta = point_2 - point_1
    theta = np.arctan2(delta.item(2), delta.item(0))
    c, s = np.cos(theta), np.sin(theta)
    Ry = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    in_xy = Ry @ delta
    Rz = _rot_z(np.arctan2(in_xy.item(1), in_xy.item(0)), 3)
    rotation = Ry.T @ Rz
    translation = rotation.T @ (point_1 + point_2) / 2
    return rotation, translation, np.linalg.norm(delta, 2)
