"""Axis helper used by the reference's demos (reference TG/path_plotter.py:3-26)."""


def set_axes_equal(ax, dimension):
    """Equal data scale on every axis: ``ax.axis('equal')`` in 2-D; in 3-D (where matplotlib has no such mode)
    all three limits are widened to the largest span about their own midpoints."""
    if dimension == 2:
        ax.axis('equal')
    if dimension == 3:
        getters = (ax.get_xlim3d, ax.get_ylim3d, ax.get_zlim3d)
        setters = (ax.set_xlim3d, ax.set_ylim3d, ax.set_zlim3d)
        limits = [g() for g in getters]
        radius = 0.5 * max(abs(hi - lo) for lo, hi in limits)
        for (lo, hi), setter in zip(limits, setters):
            mid = (lo + hi) / 2.0
            setter([mid - radius, mid + radius])
