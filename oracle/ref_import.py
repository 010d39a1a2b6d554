"""TEST INFRASTRUCTURE -- import the UNMODIFIED Python reference in place.

Only usable in the build container (``/root/reference`` does not exist on the
GPU box).  It is used by ``tests/golden/make_golden.py`` to generate the
committed fixtures and by the optional live cross-checks in ``tests/``.

The reference cannot be imported as shipped here because:
  * ``matplotlib`` and ``bsplinegenerator`` are not installed -> stub modules;
  * its four ctypes wrapper modules load
    ``constraint_functions/TrajectoryConstraintsCCode/build/src/libTrajectoryConstraints.so``
    relative to the (read-only) package (CF/turning_constraints.py:12-15) ->
    ``ctypes.CDLL`` is redirected to ``oracle/_ref/libTrajectoryConstraints.so``
    (the reference's own C++ compiled by ``oracle/build_ref.sh``).
No reference source is copied or modified.
"""
import contextlib
import ctypes
import io
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("TG_REFERENCE_ROOT", "/root/reference")
REF_LIB = os.path.join(HERE, "_ref", "libTrajectoryConstraints.so")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "trajectory_generation")) and os.path.exists(REF_LIB)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs():
    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot")
        mpl.patches = _stub("matplotlib.patches", Rectangle=object)
    if "bsplinegenerator" not in sys.modules:
        def count_number_of_control_points(control_points):
            return len(control_points) if control_points.ndim == 1 else len(control_points[0])
        b = _stub("bsplinegenerator")
        b.helper_functions = _stub("bsplinegenerator.helper_functions",
                                   count_number_of_control_points=count_number_of_control_points)
        b.bsplines = _stub("bsplinegenerator.bsplines", BsplineEvaluation=object)


_REAL_CDLL = ctypes.CDLL


class _RedirectCDLL(_REAL_CDLL):
    def __init__(self, name, *a, **k):
        if name is not None and str(name).endswith("libTrajectoryConstraints.so") and REF_ROOT in str(name):
            name = REF_LIB
        super().__init__(name, *a, **k)


def import_reference():
    """Returns the reference's ``trajectory_generation`` package (imported from
    REF_ROOT with the repo root removed from ``sys.path`` so that the drop-in
    alias package of the same name in this repo is not picked up)."""
    if not available():
        raise RuntimeError("reference not available (need %s and %s)" % (REF_ROOT, REF_LIB))
    for k in list(sys.modules):
        if k == "trajectory_generation" or k.startswith("trajectory_generation."):
            del sys.modules[k]
    _install_stubs()
    repo_root = os.path.dirname(HERE)
    saved = list(sys.path)
    sys.path[:] = [REF_ROOT] + [p for p in sys.path
                                if os.path.abspath(p or ".") not in (repo_root, REF_ROOT)]
    ctypes.CDLL = _RedirectCDLL
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import trajectory_generation.trajectory_generator as tg  # noqa: F401
            import trajectory_generation
    finally:
        ctypes.CDLL = _REAL_CDLL
        sys.path[:] = saved
    return trajectory_generation


def namespace():
    """The names a problem-definition function needs, bound to the reference."""
    import_reference()
    import importlib
    g = {}
    for mod, names in _API.items():
        m = importlib.import_module(mod)
        for n in names:
            g[n] = getattr(m, n)
    return g


_API = {
    "trajectory_generation.trajectory_generator": ["TrajectoryGenerator"],
    "trajectory_generation.constraint_data_structures.waypoint_data": ["Waypoint", "WaypointData"],
    "trajectory_generation.constraint_data_structures.dynamic_bounds": ["DerivativeBounds", "TurningBound"],
    "trajectory_generation.constraint_data_structures.obstacle": ["Obstacle"],
    "trajectory_generation.constraint_data_structures.safe_flight_corridor": [
        "SFC", "SFC_Data", "get2DRotationAndTranslationFromPoints", "get3DRotationAndTranslationFromPoints"],
    "trajectory_generation.constraint_data_structures.constraints_container": ["ConstraintsContainer"],
}
