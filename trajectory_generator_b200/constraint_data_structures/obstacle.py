"""Sphere / circle obstacles: the container the solver reads (centre, radius) plus drawing helpers kept for scripts
written against the reference (DS/obstacle.py:5-31).  Only ``Obstacle`` is on the hot path: ``pack_problem`` and
``BatchedProblem`` flatten centres coordinate-major and radii into the parameter rows read by ``tg_rows_obstacles``."""
from dataclasses import dataclass
from typing import List

import numpy as np
import numpy.typing as npt


@dataclass
class Obstacle:
    """A ball of ``radius`` around ``center`` (a d x 1 column) that the trajectory's MINVO hulls must stay out of."""
    center: npt.NDArray[np.float64]
    radius: np.double

    def as_row(self):
        """(c_0, ..., c_{d-1}, r) as a flat float64 array."""
        return np.append(np.asarray(self.center, dtype=np.float64).flatten(), float(self.radius))


class ObstacleList:
    obstacle_list: List[Obstacle]


def _draw(obstacle, ax, three_d):
    centre = np.asarray(obstacle.center, dtype=float).flatten()
    if not three_d:
        from matplotlib.patches import Circle
        ax.add_patch(Circle(tuple(centre[:2]), obstacle.radius, color="r"))
        return
    lon, lat = np.mgrid[0:2 * np.pi:20j, 0:np.pi:10j]
    unit = np.stack([np.cos(lon) * np.sin(lat), np.sin(lon) * np.sin(lat), np.cos(lat)])
    surface = centre[:3, None, None] + obstacle.radius * unit
    ax.plot_surface(surface[0], surface[1], surface[2], color="r")


def plot_2D_obstacle(obstacle: Obstacle, ax):
    _draw(obstacle, ax, False)


def plot_3D_obstacle(obstacle: Obstacle, ax):
    _draw(obstacle, ax, True)


def plot_2D_obstacles(obstacles: List[Obstacle], ax):
    for item in obstacles or ():
        _draw(item, ax, False)


def plot_3D_obstacles(obstacles: List[Obstacle], ax):
    for item in obstacles or ():
        _draw(item, ax, True)
