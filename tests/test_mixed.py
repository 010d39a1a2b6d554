"""Problems of different shapes in one call: tg_solve_mixed_host / generate_trajectories."""
import numpy as np
import pytest

import helpers
import problems


def _packed(name):
    from trajectory_generator_b200.problem import pack_problem
    d, cc, kw = problems.ALL[name](helpers.product_namespace())
    return pack_problem(d, cc, kw.get("objective_function_type", "minimal_velocity_and_time_path"),
                        kw.get("num_intervals_free_space"))


def test_mixed_entry_point_fails_loudly_without_a_device(native_lib):
    """No CPU path: without a device the call returns an error code and says why (with a device: argument checks)."""
    import torch
    from trajectory_generator_b200 import batch
    pp = _packed("obstacle2d")
    if torch.cuda.is_available():
        with pytest.raises(ValueError):
            batch.solve_mixed_host([(pp.spec, np.stack([pp.par, pp.par]), pp.x0[None])])
        return
    with pytest.raises(RuntimeError) as err:
        batch.solve_mixed_host([(pp.spec, pp.par[None], pp.x0[None])])
    assert "no CUDA device" in str(err.value)


def test_mixed_of_nothing_is_nothing(native_lib):
    from trajectory_generator_b200 import batch
    assert batch.solve_mixed_host([]) == []


@pytest.mark.gpu
def test_mixed_call_equals_bucket_by_bucket_calls(native_lib):
    """Results are those of tg_solve_host on every bucket, bit for bit (problems are independent; a problem's
    arithmetic does not depend on what else is in flight), for more buckets than concurrent workers, ragged sizes
    and an empty bucket."""
    from trajectory_generator_b200 import batch, synthetic as syn
    buckets = []
    for name, B in (("C2", 700), ("C4", 300), ("C3", 257), ("C5a", 1000), ("C2", 1), ("C5c", 33)):
        bt = syn.make(name, B)
        buckets.append((bt.spec, bt.par, bt.x0))
    bt = syn.make("C3", 4)
    buckets.insert(2, (bt.spec, bt.par[:0], bt.x0[:0]))          # an empty bucket
    mixed = batch.solve_mixed_host(buckets, jacobian="fd")
    assert len(mixed) == len(buckets)
    for (spec, par, x0), out in zip(buckets, mixed):
        if len(par) == 0:
            assert out["x"].shape[0] == 0
            continue
        ref = batch.solve_host(spec, par, x0, jacobian="fd")
        for key in ("x", "f", "status", "nit", "violation"):
            assert np.array_equal(out[key], ref[key]), key
    assert any((out["status"] == 0).any() for out in mixed)


@pytest.mark.gpu
def test_generate_trajectories_of_mixed_containers(native_lib):
    """The drop-in class on containers of different shapes (1 and 8 obstacles): each result equals the
    single-container call of the reference API."""
    from trajectory_generator_b200.trajectory_generator import TrajectoryGenerator
    ns = helpers.product_namespace()
    containers = [problems.ALL[n](ns)[1] for n in ("obstacle2d", "obstacles8", "obstacle2d", "obstacles8", "obstacles8")]
    gen = TrajectoryGenerator(2)
    results = gen.generate_trajectories(containers)
    assert len(results) == len(containers)
    assert results[0].control_points.shape == results[1].control_points.shape       # same N, different row counts
    for cc, res in zip(containers, results):
        cps, scale, viol = gen.generate_trajectory(cc)
        assert np.array_equal(cps, res.control_points) and scale == res.scale_factor and viol == res.is_violation
        assert res.status == gen.last_result.status and res.nit == gen.last_result.nit
