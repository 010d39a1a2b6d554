"""Sphere / circle obstacles (reference DS/obstacle.py:5-31)."""
from dataclasses import dataclass
import numpy as np
import numpy.typing as npt


@dataclass
class Obstacle:
    center: npt.NDArray[np.float64]
    radius: np.double


class ObstacleList:
    obstacle_list: 'list[Obstacle]'


def plot_2D_obstacle(obstacle: Obstacle, ax):
    import matplotlib.pyplot as plt
    ax.add_patch(plt.Circle((obstacle.center.item(0), obstacle.center.item(1)), obstacle.radius, color='r'))


def plot_2D_obstacles(obstacles: 'list[Obstacle]', ax):
    for obstacle in obstacles:
        plot_2D_obstacle(obstacle, ax)


def plot_3D_obstacle(obstacle: Obstacle, ax):
    u, v = np.mgrid[0:2 * np.pi:20j, 0:np.pi:10j]
    r, c = obstacle.radius, obstacle.center
    ax.plot_surface(r * np.cos(u) * np.sin(v) + c.item(0), r * np.sin(u) * np.sin(v) + c.item(1),
                    r * np.cos(v) + c.item(2), color="r")


def plot_3D_obstacles(obstacles: 'list[Obstacle]', ax):
    for obstacle in obstacles:
        plot_3D_obstacle(obstacle, ax)
