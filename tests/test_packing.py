"""Host logic: ConstraintsContainer -> packed problem (descriptor, parameter row, x0, bounds) against the
fixtures recorded from the reference, and the vectorised synthetic generators against the packer."""
import numpy as np
import pytest

import helpers
import problems


@pytest.mark.parametrize("name", list(problems.ALL))
def test_pack_matches_reference_fixture(native_lib, name):
    from trajectory_generator_b200.problem import pack_problem
    G = helpers.load_golden()["problems"][name]
    d, cc, kw = problems.ALL[name](helpers.product_namespace())
    pp = pack_problem(d, cc, kw.get("objective_function_type", "minimal_velocity_and_time_path"),
                      kw.get("num_intervals_free_space"))
    L = pp.layout
    assert (L.d, L.N, L.n, L.m, L.meq) == (G["dimension"], G["N"], G["n"], G["m"], G["meq"])
    assert np.array_equal(np.clip(pp.x0, pp.xl, pp.xu), np.array(G["x0"]))
    assert np.array_equal(pp.xl, np.array(G["xl"])) and np.array_equal(pp.xu, np.array(G["xu"]))


def test_invalid_inputs_raise_like_the_reference(native_lib):
    from trajectory_generator_b200.problem import pack_problem
    ns = helpers.product_namespace()
    d, cc, kw = problems.obstacle2d(ns)
    with pytest.raises(Exception, match="Invalid objective function type"):
        pack_problem(d, cc, "fastest_path")
    # location-only terminal waypoint: the reference dies with IndexError inside scipy
    W, WD = ns["Waypoint"], ns["WaypointData"]
    col = lambda *v: np.array([[float(x)] for x in v])
    cc2 = ns["ConstraintsContainer"](WD((W(location=col(0, 0)), W(location=col(1, 1), velocity=col(1, 0)))))
    with pytest.raises(IndexError):
        pack_problem(2, cc2)
    with pytest.raises(Exception, match="Bound type"):
        ns["TurningBound"](1.0, "yaw")
    with pytest.raises(Exception, match="general max velocity"):
        ns["DerivativeBounds"](max_upward_velocity=1.0)


@pytest.mark.parametrize("name", ["C2", "C3", "C4", "C5a", "C5c"])
def test_synthetic_generators_match_the_packer(native_lib, name):
    from trajectory_generator_b200 import synthetic as syn
    from trajectory_generator_b200.problem import pack_problem
    b = syn.make(name, 257)
    expected_shape = {"C2": (17, 19, 8, 35), "C3": (37, 18, 16, 18), "C4": (34, 209, 15, 71), "C5a": (17, 11, 8, 11),
                      "C5c": (17, 11, 8, 11)}[name]        # (n, m, meq, P) of SURVEY.md 8 table
    L = b.layout
    assert (L.n, L.m, L.meq, L.P) == expected_shape
    for i in (0, 3, 256):
        d, cc, kw = syn.container_for(b, i)
        pp = pack_problem(d, cc, kw.get("objective_function_type", syn.OBJECTIVE[name]), kw.get("num_intervals_free_space"))
        assert np.array_equal(pp.spec, b.spec)
        assert np.abs(pp.par - b.par[i]).max() <= 1e-13
        assert np.abs(pp.x0 - b.x0[i]).max() <= 1e-13
    # same seed -> same batch (the CPU baseline regenerates it in its worker processes)
    assert np.array_equal(syn.make(name, 257).par, b.par)


def _rows_from_container(d, cc, kw):
    """Arguments of batched.assemble_rows for ONE container (a batch of two identical rows), read off the
    dataclasses the way a user of the batched front end would."""
    import torch
    wd, db, tb, obs = cc.waypoint_constraints, cc.derivative_constraints, cc.turning_constraint, cc.obstacle_constraints
    sw, ew = wd.start_waypoint, wd.end_waypoint
    t = lambda a: None if a is None else torch.from_numpy(np.stack([np.asarray(a, dtype=np.float64).reshape(-1)] * 2))
    args = dict(start=t(sw.location), end=t(ew.location),
                start_zero_velocity=sw.checkIfZeroVel(), end_zero_velocity=ew.checkIfZeroVel(),
                start_velocity=None if sw.checkIfZeroVel() else t(sw.velocity),
                end_velocity=None if ew.checkIfZeroVel() else t(ew.velocity),
                end_is_target=bool(ew.is_target), start_direction=t(sw.direction), end_direction=t(ew.direction),
                start_acceleration=t(sw.acceleration), end_acceleration=t(ew.acceleration),
                objective_function_type=kw.get("objective_function_type", "minimal_velocity_and_time_path"),
                num_intervals_free_space=kw.get("num_intervals_free_space"))
    if wd.intermediate_locations is not None:
        args["intermediate_locations"] = torch.from_numpy(np.stack([wd.intermediate_locations] * 2))
        if wd.intermediate_velocities is not None:
            args["intermediate_velocities"] = torch.from_numpy(np.stack([wd.intermediate_velocities] * 2))
    if db is not None:
        args.update(max_velocity=db.max_velocity, max_acceleration=db.max_acceleration, max_jerk=db.max_jerk,
                    gravity=db.gravity, max_upward_velocity=db.max_upward_velocity,
                    max_horizontal_velocity=db.max_horizontal_velocity, min_velocity=db.min_velocity)
        if db.checkIfTangentialAccelerationActive():
            args["tangential_acceleration"] = (db.min_tangential_acceleration, db.max_tangential_acceleration)
    if tb is not None and tb.checkIfTurningBoundActive():
        args["turning"] = (tb.bound_type, tb.max_turning_bound)
    if obs is not None:
        ctr = np.array([[float(np.asarray(o.center).flatten()[c]) for c in range(d)] for o in obs])      # [K, d]
        args["obstacle_centers"] = torch.from_numpy(np.stack([ctr] * 2))
        args["obstacle_radii"] = torch.from_numpy(np.stack([np.array([float(o.radius) for o in obs])] * 2))
    return args


@pytest.mark.parametrize("name", [n for n in problems.ALL if "sfc" not in n])
def test_batched_assembly_equals_the_packer(native_lib, name):
    """The batched front end's descriptor and parameter rows (assembled with torch ops on any device) are those of
    pack_problem, for every field of a container: directions, accelerations, zero-velocity and target waypoints,
    every derivative bound, tangential acceleration, all turning kinds, obstacles, intermediate waypoints.
    (Corridor blocks are filled by a CUDA kernel: tests/test_batched.py, tests/test_build.py.)"""
    import torch
    from trajectory_generator_b200.batched import assemble_rows
    from trajectory_generator_b200.problem import pack_problem
    d, cc, kw = problems.ALL[name](helpers.product_namespace())
    pp = pack_problem(d, cc, kw.get("objective_function_type", "minimal_velocity_and_time_path"),
                      kw.get("num_intervals_free_space"))
    spec, blocks, i_sfc, ipc = assemble_rows(d, **_rows_from_container(d, cc, kw))
    assert i_sfc is None and ipc is None
    assert np.array_equal(spec, pp.spec), (spec, pp.spec)
    par = torch.cat(blocks, 1).numpy()
    assert par.shape == (2, pp.layout.P)
    assert np.array_equal(par[0], pp.par) and np.array_equal(par[1], pp.par)


def test_batched_assembly_rejects_what_the_dataclasses_reject(native_lib):
    import torch
    from trajectory_generator_b200.batched import assemble_rows
    z = torch.zeros((2, 2), dtype=torch.float64)
    with pytest.raises(IndexError):
        assemble_rows(2, z, z + 1)                                              # location-only terminal waypoint
    with pytest.raises(Exception, match="Invalid objective function type"):
        assemble_rows(2, z, z + 1, z + 1, z + 1, objective_function_type="fastest_path")
    with pytest.raises(Exception, match="general max velocity"):
        assemble_rows(2, z, z + 1, z + 1, z + 1, max_upward_velocity=1.0)
    with pytest.raises(Exception, match="non-zero velocity"):
        assemble_rows(2, z, z + 1, z + 1, z + 1, start_direction=z + 1)


def test_batched_assembly_equals_the_packer_on_random_containers(native_lib):
    """Random combinations of every container field (seeded): assemble_rows == pack_problem, descriptor and row."""
    import torch
    from trajectory_generator_b200.batched import assemble_rows
    from trajectory_generator_b200.problem import pack_problem
    ns = helpers.product_namespace()
    W, WD, DB, TB, Ob, CC = (ns[k] for k in ("Waypoint", "WaypointData", "DerivativeBounds", "TurningBound", "Obstacle",
                                               "ConstraintsContainer"))
    rng = np.random.default_rng(20261018)
    objectives = ["minimal_time_path", "minimal_distance_path", "minimal_velocity_path", "minimal_acceleration_path",
                  "minimal_distance_and_time_path", "minimal_velocity_and_time_path", "minimal_acceleration_and_time_path",
                  "minimal_time_path_velocity_penalty"]
    checked = 0
    for trial in range(120):
        d = int(rng.integers(2, 4))
        vec = lambda: rng.normal(size=(d, 1)) * 3

        def terminal(is_end):
            kind = rng.integers(0, 5)
            kwargs = dict(location=vec())
            if kind == 0:
                kwargs["velocity"] = vec()
            elif kind == 1:
                kwargs["velocity"] = np.zeros((d, 1))                       # zero-velocity waypoint
            elif kind == 2:
                kwargs["velocity"] = np.zeros((d, 1)); kwargs["direction"] = vec()
            elif kind == 3:
                kwargs["direction"] = vec()
            else:
                kwargs["velocity"] = vec(); kwargs["acceleration"] = vec()
            if rng.random() < 0.2 and kind != 0:
                kwargs["acceleration"] = vec()
            if is_end and kind in (0, 4) and rng.random() < 0.3:
                kwargs["is_target"] = True
            return W(**kwargs)
        wps = [terminal(False)]
        niw = int(rng.integers(0, 3))
        with_iv = rng.random() < 0.5
        for _ in range(niw):
            wps.append(W(location=vec(), velocity=vec() if with_iv else None))
        wps.append(terminal(True))
        db = None
        if rng.random() < 0.8:
            f = {}
            if rng.random() < 0.7: f["max_velocity"] = float(rng.uniform(3, 9))
            if rng.random() < 0.5: f["max_acceleration"] = float(rng.uniform(1, 9))
            if rng.random() < 0.3: f["max_jerk"] = float(rng.uniform(1, 9))
            if rng.random() < 0.3: f["gravity"] = float(rng.uniform(0.1, 1))
            if rng.random() < 0.3: f["min_velocity"] = float(rng.uniform(0.01, 0.5))
            if "max_velocity" in f and rng.random() < 0.4: f["max_upward_velocity"] = f["max_velocity"] * 0.5
            if "max_velocity" in f and rng.random() < 0.4: f["max_horizontal_velocity"] = f["max_velocity"] * 0.8
            if rng.random() < 0.3:
                f["min_tangential_acceleration"] = -float(rng.uniform(1, 5)); f["max_tangential_acceleration"] = float(rng.uniform(1, 5))
            db = DB(**f)
        tb = TB(float(rng.uniform(0.5, 5)), ["curvature", "angular_rate", "centripetal_acceleration"][rng.integers(0, 3)]) \
            if rng.random() < 0.6 else None
        obs = [Ob(center=vec(), radius=float(rng.uniform(0.2, 1))) for _ in range(int(rng.integers(1, 4)))] \
            if rng.random() < 0.5 else None
        cc = CC(WD(tuple(wps)), db, tb, None, obs)
        kw = dict(objective_function_type=objectives[rng.integers(0, len(objectives))])
        if rng.random() < 0.5:
            kw["num_intervals_free_space"] = int(rng.integers(4, 12))
        try:
            pp = pack_problem(d, cc, kw["objective_function_type"], kw.get("num_intervals_free_space"))
        except IndexError:
            with pytest.raises(IndexError):         # a terminal waypoint without any derivative row: both refuse
                assemble_rows(d, **_rows_from_container(d, cc, kw))
            continue
        spec, blocks, i_sfc, ipc = assemble_rows(d, **_rows_from_container(d, cc, kw))
        assert np.array_equal(spec, pp.spec), (trial, spec, pp.spec)
        par = torch.cat(blocks, 1).numpy()
        assert par.shape == (2, pp.layout.P) and np.array_equal(par[0], pp.par), trial
        checked += 1
    assert checked >= 80


def test_batched_problem_constructor_plumbing(native_lib, monkeypatch):
    """BatchedProblem's constructor end to end on the CPU with the two CUDA builders stubbed out (tensors that claim
    to be CUDA tensors): descriptor, parameter rows, the arguments handed to the builders.  The builders themselves
    and the solve are covered on the GPU (tests/test_batched.py, tests/test_build.py)."""
    import torch
    from trajectory_generator_b200 import batched, builder, synthetic as syn

    class Claimed(torch.Tensor):
        @property
        def is_cuda(self):
            return True

    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).as_subclass(Claimed)
    calls = {}

    def fake_guess(spec, seq, wseq, scale):
        calls["guess"] = (np.array(spec), tuple(seq.shape), None if wseq is None else tuple(wseq.shape), scale)
        return torch.zeros((seq.shape[0], batched.pk.Layout(spec).n), dtype=torch.float64)

    def fake_boxes(spec, points, pads, par):
        calls["boxes"] = (tuple(points.shape), tuple(pads.shape), tuple(par.shape))

    monkeypatch.setattr(builder, "initial_guess_batch", fake_guess)
    monkeypatch.setattr(builder, "sfc_boxes_batch", fake_boxes)
    # C2: obstacles; C3: intermediate waypoints; C4: corridors (as tests/test_batched.py builds them)
    b = syn.make("C2", 8); r = b.raw
    bp = batched.BatchedProblem(2, t(r["start"]), t(r["goal"]), t(r["v0"]), t(r["v1"]), max_velocity=r["vmax"],
                                max_acceleration=r["amax"], turning=("angular_rate", r["turn"]),
                                obstacle_centers=t(r["centers"]), obstacle_radii=t(r["radii"]))
    assert np.array_equal(bp.spec, b.spec) and np.array_equal(bp.par.numpy(), b.par)
    assert calls["guess"][1:] == ((8, 2, 2), None, 1.0) and (bp.B, bp.d, bp.N) == (8, 2, b.layout.N)
    b = syn.make("C3", 8); p, v = b.raw["points"], b.raw["velocities"]
    bp = batched.BatchedProblem(2, t(p[:, :, 0]), t(p[:, :, 3]), t(v[:, :, 0]), t(v[:, :, 3]),
                                intermediate_locations=t(p[:, :, 1:3]), intermediate_velocities=t(v[:, :, 1:3]),
                                max_velocity=b.raw["vmax"], turning=("curvature", b.raw["turn"]), num_intervals_free_space=14)
    assert np.array_equal(bp.spec, b.spec) and np.array_equal(bp.par.numpy(), b.par)
    assert calls["guess"][1:3] == ((8, 2, 4), (8, 2, 4))
    b = syn.make("C4", 8); p = b.raw["points"]
    pad = b.raw["dims"].copy(); pad[:, :, 0] -= np.linalg.norm(p[:, :, 1:] - p[:, :, :-1], 2, 1)
    bp = batched.BatchedProblem(3, t(p[:, :, 0]), t(p[:, :, 4]), t(b.raw["v0"]), end_zero_velocity=True,
                                max_velocity=b.raw["vmax"], max_acceleration=b.raw["amax"], corridor_points=t(p),
                                corridor_pads=t(pad), objective_function_type="minimal_velocity_path")
    assert np.array_equal(bp.spec, b.spec) and bp.par.shape == b.par.shape
    assert calls["boxes"] == ((8, 3, 5), (8, 4, 3), (8, b.layout.P)) and calls["guess"][1:3] == ((8, 3, 5), None)
    # a field of the extended set goes through the constructor's keyword pass-through
    bp = batched.BatchedProblem(2, t(r["start"]), t(r["goal"]), None, t(r["v1"]), start_direction=t(r["v0"]), max_jerk=3.0)
    assert bp.spec[batched.pk.SP_START_DIR] == 1 and bp.spec[batched.pk.SP_DB_JERK] == 1 and bp.layout.nws == 1


def test_fixed_shape_descriptors_are_the_synthetic_configurations(native_lib):
    """csrc/tg_shape.h lists the BASELINE shapes whose kernels are instantiated with a compile-time descriptor:
    each must equal the descriptor synthetic.py (and pack_problem) builds for that configuration, and other shapes
    must fall through to the generic kernels."""
    import ctypes
    from trajectory_generator_b200 import synthetic as syn
    import problems
    f = native_lib.tg_fixed_shape_index
    f.argtypes = [ctypes.POINTER(ctypes.c_int)]; f.restype = ctypes.c_int
    got = {}
    for name in syn.CONFIGS:
        spec = np.ascontiguousarray(syn.make(name, 2).spec, dtype=np.int32)
        got[name] = f(spec.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    assert got == {"C2": 1, "C3": 2, "C4": 3, "C5a": 4, "C5c": 5}
    from trajectory_generator_b200.problem import pack_problem
    d, cc, kw = problems.sfc3d(helpers.product_namespace())           # shipped 3-corridor problem: not a fixed shape
    pp = pack_problem(d, cc, kw["objective_function_type"])
    assert f(np.ascontiguousarray(pp.spec, dtype=np.int32).ctypes.data_as(ctypes.POINTER(ctypes.c_int))) == 0
    d, cc, kw = problems.obstacles8(helpers.product_namespace())      # the C2 shape through the drop-in packer
    pp = pack_problem(d, cc)
    assert f(np.ascontiguousarray(pp.spec, dtype=np.int32).ctypes.data_as(ctypes.POINTER(ctypes.c_int))) == 1


def test_vectorised_packing_equals_per_container_packing(native_lib):
    """problem.pack_problems (grouped by shape, one gather per field) against pack_problem container by container:
    every fixture problem, and the synthetic batches (where the rows must also equal the vectorised generators')."""
    from trajectory_generator_b200 import synthetic as syn
    from trajectory_generator_b200.problem import pack_problem, pack_problems
    ns = helpers.product_namespace()
    by_args = {}
    for name, make in problems.ALL.items():
        d, cc, kw = make(ns)
        by_args.setdefault((d, kw.get("objective_function_type", "minimal_velocity_and_time_path"),
                            kw.get("num_intervals_free_space")), []).append((name, cc))
    for (d, obj, nifs), items in by_args.items():
        groups = pack_problems(d, [cc for _, cc in items], obj, nifs)
        assert sorted(int(i) for g in groups for i in g.indices) == list(range(len(items)))
        for g in groups:
            for row, i in enumerate(g.indices):
                name, cc = items[int(i)]
                pp = pack_problem(d, cc, obj, nifs)
                assert np.array_equal(g.spec, pp.spec), name
                assert np.array_equal(g.par[row], pp.par), name
                assert np.array_equal(g.x0[row], np.clip(pp.x0, pp.xl, pp.xu)), name
                assert np.array_equal(g.xl, pp.xl) and np.array_equal(g.xu, pp.xu), name
    for cfg in ("C2", "C3", "C4", "C5a"):
        b = syn.make(cfg, 48)
        items = [syn.container_for(b, i) for i in range(48)]
        d, kw = items[0][0], items[0][2]
        groups = pack_problems(d, [cc for _, cc, _ in items], kw.get("objective_function_type", syn.OBJECTIVE[cfg]),
                               kw.get("num_intervals_free_space"))
        assert len(groups) == 1 and np.array_equal(groups[0].indices, np.arange(48))
        assert np.array_equal(groups[0].spec, b.spec)
        assert np.abs(groups[0].par - b.par).max() <= 1e-12 * max(1.0, np.abs(b.par).max())
        assert np.abs(groups[0].x0 - b.x0).max() <= 1e-12 * max(1.0, np.abs(b.x0).max())
