"""Batched front end (trajectory_generator_b200/batched.py): the arrays it assembles on the GPU equal the vectorised
host generators (which tests/test_packing.py ties to pack_problem, i.e. to the per-container path)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _build(name, b):
    from trajectory_generator_b200.batched import BatchedProblem
    r = b.raw
    if name == "C2":
        return BatchedProblem(2, _t(r["start"]), _t(r["goal"]), _t(r["v0"]), _t(r["v1"]), max_velocity=r["vmax"],
                              max_acceleration=r["amax"], turning=("angular_rate", r["turn"]),
                              obstacle_centers=_t(r["centers"]), obstacle_radii=_t(r["radii"]))
    if name == "C3":
        p, v = r["points"], r["velocities"]
        return BatchedProblem(2, _t(p[:, :, 0]), _t(p[:, :, 3]), _t(v[:, :, 0]), _t(v[:, :, 3]),
                              intermediate_locations=_t(p[:, :, 1:3]), intermediate_velocities=_t(v[:, :, 1:3]),
                              max_velocity=r["vmax"], turning=("curvature", r["turn"]), num_intervals_free_space=14)
    p = r["points"]
    seglen = np.linalg.norm(p[:, :, 1:] - p[:, :, :-1], 2, 1)
    pad = r["dims"].copy(); pad[:, :, 0] -= seglen
    return BatchedProblem(3, _t(p[:, :, 0]), _t(p[:, :, 4]), _t(r["v0"]), end_zero_velocity=True, max_velocity=r["vmax"],
                          max_acceleration=r["amax"], corridor_points=_t(p), corridor_pads=_t(pad),
                          objective_function_type="minimal_velocity_path")


@pytest.mark.parametrize("name", ["C2", "C3", "C4"])
def test_batched_problem_equals_host_generator(native_lib, name):
    from trajectory_generator_b200 import synthetic as syn
    b = syn.make(name, 256)
    bp = _build(name, b)
    assert np.array_equal(bp.spec, b.spec)
    par = bp.par.cpu().numpy(); x0 = bp.x0.cpu().numpy()
    assert par.shape == b.par.shape and x0.shape == b.x0.shape
    assert np.abs(par - b.par).max() <= 1e-12 * max(1.0, np.abs(b.par).max())
    assert np.abs(x0 - b.x0).max() <= 1e-12 * max(1.0, np.abs(b.x0).max())


def test_batched_problem_solves_and_samples(native_lib):
    """End to end on device tensors: build -> solve -> sample; the answers are those of the array-level API."""
    from trajectory_generator_b200 import batch, synthetic as syn
    import tg_oracle_sampling as osamp
    b = syn.make("C4", 256)
    bp = _build("C4", b)
    out = bp.solve()
    ref = batch.solve_host(b.spec, b.par, b.x0, jacobian="fd")
    st = out["status"].cpu().numpy()
    assert (st == 0).all() and (ref["status"] == 0).all()
    # inputs agree to ~1e-16 (device vs numpy arithmetic of the initial guess / boxes); forward differences amplify that
    dx = np.abs(out["x"].cpu().numpy() - ref["x"]).max(1)
    # (a few problems are ones the reference itself moves by 1e-4 ... 1e-2 from x0 +- 1 ulp)
    assert (dx <= 1e-5).mean() >= 0.97 and dx.max() <= 1e-2
    cps, scale = bp.control_points(out)
    pos = bp.sample(out, num_points=50).cpu().numpy()
    for i in (0, 17, 255):
        assert np.abs(pos[i] - osamp.dataset(cps[i].cpu().numpy(), 50)).max() <= 1e-12 * 20
