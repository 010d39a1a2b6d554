"""Output sampling: the numpy oracle against fixtures recorded from the unmodified reference
(CPU), and the CUDA samplers against the same fixtures through the reference's function names (GPU)."""
import numpy as np
import pytest

import helpers


def _golden():
    return helpers.load_golden("sampling.json")


def test_oracle_matches_reference_fixtures():
    import tg_oracle_sampling as osamp
    G = _golden()
    for name, rec in G["splines"].items():
        cps = np.array(rec["control_points"]); sc = rec["scale_factor"]
        for num, ref in rec["dataset"].items():
            assert np.array_equal(osamp.dataset(cps, int(num)), np.array(ref)), (name, num)
        for r, ref in rec["derivative_dataset"].items():
            assert np.array_equal(osamp.derivative_dataset(cps, int(r), sc, 50), np.array(ref)), (name, r)
        for case in rec["discrete"]:
            d, t, rem, end = osamp.discrete_steps(cps, case["start_time"], case["offset"], case["dt"], sc, case["r"])
            assert np.array_equal(d, np.array(case["data"])) and np.array_equal(t, np.array(case["time"]))
            assert rem == case["remainder"] and end == case["end"]
        assert np.array_equal(osamp.timedataset(cps, np.array(rec["timedataset"]["time"]), sc), np.array(rec["timedataset"]["data"]))
    lst = [np.array(G["splines"][k]["control_points"]) for k in ("concat1", "concat2", "concat3")]
    sc = [G["splines"][k]["scale_factor"] for k in ("concat1", "concat2", "concat3")]
    for case in G["concatenate"]:
        d, t = osamp.concatenate(2, case["dt"], case["start_time"], lst, sc, case["r"])
        assert np.array_equal(d, np.array(case["data"])) and np.array_equal(t, np.array(case["time"]))


def test_oracle_matches_reference_fixtures_of_other_orders_and_helpers():
    """orders 2, 4, 5 (TG/matrix_evaluation.py:197-262) and the single-point helpers (:183-232)"""
    import tg_oracle_sampling as osamp
    from trajectory_generator_b200 import matrix_evaluation as me
    G = _golden()
    for order, rec in G["orders"].items():
        order = int(order)
        cps = np.array(rec["control_points"]); sc = rec["scale_factor"]
        assert np.array_equal(osamp.dataset(cps, 41, order), np.array(rec["dataset"]))
        for r, ref in rec["derivative_dataset"].items():
            assert np.array_equal(osamp.derivative_dataset(cps, int(r), sc, 37, order), np.array(ref)), (order, r)
        for case in rec["discrete"]:
            d, t, rem, end = osamp.discrete_steps(cps, case["start_time"], case["offset"], case["dt"], sc, case["r"], order)
            assert np.array_equal(d, np.array(case["data"])) and np.array_equal(t, np.array(case["time"]))
    for o, M in G["helpers"]["M"].items():
        assert np.array_equal(me.get_M_matrix(int(o)), np.array(M))            # host-side constants of the drop-in module
    for case in G["helpers"]["points"]:
        cp = np.array(case["control_points"])
        got = osamp.point_on_interval(cp, case["t"], case["tj"], case["scale"], case["r"])
        assert np.allclose(got, np.array(case["value"]), rtol=1e-13, atol=1e-13)
        order = cp.shape[1] - 1
        T = (me.get_T_vector(order, case["t"], case["tj"], case["scale"]) if case["r"] == 0 else
             me.get_T_derivative_vector(order, case["t"], case["tj"], case["r"], case["scale"]))
        assert np.array_equal(T, np.array(case["T"]))
    with pytest.raises(Exception, match="Cannot return M matrix"):
        me.get_M_matrix(1)                                                       # the reference's if / elif fall-through


def _close(a, b):
    a = np.asarray(a, dtype=float); b = np.asarray(b, dtype=float)
    assert a.shape == b.shape, (a.shape, b.shape)
    # same formulas; pow() vs products and the order of two 4-term sums may differ in the last place
    assert np.abs(a - b).max() <= 1e-12 * max(1.0, np.abs(b).max())


@pytest.mark.gpu
def test_cuda_samplers_match_reference_fixtures(native_lib):
    import trajectory_generation.matrix_evaluation as me
    G = _golden()
    for name, rec in G["splines"].items():
        cps = np.array(rec["control_points"]); sc = rec["scale_factor"]
        for num, ref in rec["dataset"].items():
            _close(me.matrix_bspline_evaluation_for_dataset(3, cps, int(num)), ref)
        for r, ref in rec["derivative_dataset"].items():
            _close(me.matrix_bspline_derivative_evaluation_for_dataset(3, int(r), sc, cps, 50), ref)
        for case in rec["discrete"]:
            if case["r"] == 0:
                d, t, rem, end = me.matrix_bspline_evaluation_for_discrete_steps(3, cps, case["start_time"], case["offset"], case["dt"], sc)
            else:
                d, t, rem, end = me.matrix_bspline_derivative_evaluation_for_discrete_steps(3, case["r"], sc, cps, case["start_time"], case["offset"], case["dt"])
            _close(d, case["data"])
            assert np.array_equal(t, np.array(case["time"]))          # the clock is numpy.linspace bit for bit
            assert rem == case["remainder"] and end == case["end"]
        _close(me.matrix_bspline_evaluation_for_timedataset(3, cps, np.array(rec["timedataset"]["time"]), sc), rec["timedataset"]["data"])


@pytest.mark.gpu
def test_cuda_concatenater_matches_reference_fixture(native_lib):
    from trajectory_generation.spline_data_concatenater import SplineDataConcatenater
    G = _golden()
    lst = [np.array(G["splines"][k]["control_points"]) for k in ("concat1", "concat2", "concat3")]
    sc = [G["splines"][k]["scale_factor"] for k in ("concat1", "concat2", "concat3")]
    for case in G["concatenate"]:
        d, t = SplineDataConcatenater(2).concatenate_spline_data(case["dt"], case["start_time"], [3, 3, 3], lst, sc, derivative_order=case["r"])
        _close(d, case["data"])
        assert np.array_equal(t, np.array(case["time"]))


@pytest.mark.gpu
def test_batched_sampling_of_solver_rows(native_lib):
    """Batched form on the solver's own variable rows x[B][n] (control points | scale | ...): equals the oracle per
    trajectory, positions and velocities, uniform and discrete clocks with ragged sample counts."""
    import torch
    import tg_oracle_sampling as osamp
    from trajectory_generator_b200 import matrix_evaluation as me, synthetic as syn
    b = syn.make("C2", 64)
    L = b.layout
    rows = torch.from_numpy(b.x0).cuda()
    rows[:, L.ia] = torch.linspace(0.5, 2.0, 64, dtype=torch.float64)
    pos = me.sample_batch((rows, L.d, L.N), num_points=33).cpu().numpy()
    vel = me.sample_batch((rows, L.d, L.N), derivative_order=1, num_points=33).cpu().numpy()
    data, times, counts = me.sample_batch((rows, L.d, L.N), dt=0.37)
    data, times, counts = data.cpu().numpy(), times.cpu().numpy(), counts.cpu().numpy()
    x = rows.cpu().numpy()
    assert len(set(counts.tolist())) > 1
    for i in range(64):
        cps = x[i, :L.d * L.N].reshape(L.d, L.N); sc = x[i, L.ia]
        _close(pos[i], osamp.dataset(cps, 33))
        _close(vel[i], osamp.derivative_dataset(cps, 1, sc, 33))
        d, t, rem, end = osamp.discrete_steps(cps, 0.0, 0.0, 0.37, sc)
        assert counts[i] == d.shape[1]
        _close(data[i, :, :counts[i]], d)
        assert np.array_equal(times[i, :counts[i]], t)


@pytest.mark.gpu
def test_cuda_samplers_of_other_orders_and_point_helpers(native_lib):
    """orders 2, 4, 5 through the reference's function names, and evaluate_point_[derivative_]on_interval (a5), against
    fixtures recorded from the unmodified reference"""
    import torch
    import trajectory_generation.matrix_evaluation as me
    G = _golden()
    for order, rec in G["orders"].items():
        order = int(order)
        cps = np.array(rec["control_points"]); sc = rec["scale_factor"]
        _close(me.matrix_bspline_evaluation_for_dataset(order, cps, 41), rec["dataset"])
        for r, ref in rec["derivative_dataset"].items():
            _close(me.matrix_bspline_derivative_evaluation_for_dataset(order, int(r), sc, cps, 37), ref)
        for case in rec["discrete"]:
            if case["r"] == 0:
                d, t, rem, end = me.matrix_bspline_evaluation_for_discrete_steps(order, cps, case["start_time"], case["offset"], case["dt"], sc)
            else:
                d, t, rem, end = me.matrix_bspline_derivative_evaluation_for_discrete_steps(order, case["r"], sc, cps, case["start_time"], case["offset"], case["dt"])
            _close(d, case["data"])
            assert np.array_equal(t, np.array(case["time"]))
            assert rem == case["remainder"] and end == case["end"]
    pts = G["helpers"]["points"]
    for case in pts:
        cp = np.array(case["control_points"])
        got = (me.evaluate_point_on_interval(cp, case["t"], case["tj"], case["scale"]) if case["r"] == 0 else
               me.evaluate_point_derivative_on_interval(cp, case["t"], case["tj"], case["scale"], case["r"]))
        assert got.shape == (3, 1)
        _close(got, case["value"])
    # batched form: every order-3 case in one launch
    from trajectory_generator_b200 import matrix_evaluation as tme
    c3 = [c for c in pts if len(c["control_points"][0]) == 4 and c["r"] == 0]
    t = lambda key: torch.tensor([c[key] for c in c3], dtype=torch.float64, device="cuda")
    out = tme.interval_points_batch(torch.tensor([c["control_points"] for c in c3], dtype=torch.float64, device="cuda"),
                                    t("t"), t("tj"), t("scale")).cpu().numpy()
    for row, c in zip(out, c3):
        _close(row[:, None], c["value"])
    with pytest.raises(Exception, match="Cannot return M matrix"):
        me.matrix_bspline_evaluation_for_dataset(1, np.zeros((2, 5)), 10)
