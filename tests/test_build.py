"""Problem construction: the numpy oracle against fixtures recorded from the unmodified reference
(CPU), and the CUDA builders against the same fixtures and against the vectorised host generators (GPU)."""
import numpy as np
import pytest

import helpers


def _golden():
    return helpers.load_golden("build.json")


def test_oracle_matches_reference_fixtures():
    import tg_oracle_build as ob
    G = _golden()
    for rec in G["initial"]:
        seq = np.array(rec["seq"])
        x0 = ob.initial_variables(rec["N"], seq, rec["d"], 0, seq if rec["niw"] else None)
        assert np.array_equal(x0, np.array(rec["x0"])), rec["N"]
    for rec in G["boxes"]:
        R, lo, hi, ln = ob.box_from_points(rec["p1"], rec["p2"], rec["pad"])
        assert np.array_equal(R, np.array(rec["R"])) and np.array_equal(lo, np.array(rec["lower"]))
        assert np.array_equal(hi, np.array(rec["upper"])) and ln == rec["length"]
    for rec in G["intervals"]:
        assert ob.intervals_per_corridor(rec["points"], rec["min"]) == rec["ipc"]
        assert sum(rec["ipc"]) == rec["num_intervals"]


def _spec(d, N, niw=0, ncorr=0):
    from trajectory_generator_b200 import problem as pk
    spec = np.zeros(pk.SP_COUNT, dtype=np.int32)
    spec[pk.SP_DIM], spec[pk.SP_NCP] = d, N
    spec[pk.SP_OBJECTIVE] = pk.OBJECTIVES.index("minimal_velocity_and_time_path")
    spec[pk.SP_START_VEL] = spec[pk.SP_END_VEL] = 1
    spec[pk.SP_NIW] = niw
    spec[pk.SP_IW_VEL] = 1 if niw else 0
    if ncorr:
        spec[pk.SP_NCORR] = ncorr
        spec[pk.SP_IPC0:pk.SP_IPC0 + ncorr] = (N - 3) // ncorr
    return spec


@pytest.mark.gpu
def test_cuda_initial_guess_matches_reference_fixtures(native_lib):
    import torch
    from trajectory_generator_b200 import builder
    for rec in _golden()["initial"]:
        seq = np.array(rec["seq"])
        spec = _spec(rec["d"], rec["N"], rec["niw"])
        t = torch.from_numpy(seq[None]).cuda()
        x0 = builder.initial_guess_batch(spec, t, t if rec["niw"] else None)[0].cpu().numpy()
        ref = np.array(rec["x0"])
        assert x0.shape == ref.shape
        k = rec["d"] * rec["N"]
        if seq.shape[1] == 2:
            # straight line = numpy.linspace, reproduced bit for bit
            assert np.array_equal(x0[:k + 1], ref[:k + 1]), (rec["d"], rec["N"])
        else:
            # the walk uses numpy's arithmetic (no contraction), so its branches fall as the reference's; the unit
            # vectors divide by a BLAS dot-product norm in the reference, which may differ in the last place
            assert np.abs(x0[:k + 1] - ref[:k + 1]).max() <= 1e-13 * max(1.0, np.abs(ref[:k]).max()), (rec["d"], rec["N"])
        assert np.abs(x0[k + 1:] - ref[k + 1:]).max(initial=0.0) <= 1e-13 * (rec["N"] - 3)


@pytest.mark.gpu
def test_cuda_boxes_match_reference_fixtures(native_lib):
    import torch
    from trajectory_generator_b200 import builder
    from trajectory_generator_b200.problem import Layout
    for rec in _golden()["boxes"]:
        d = rec["d"]
        spec = _spec(d, 4, 0, 1)
        lay = Layout(spec)
        pts = torch.tensor(np.stack([rec["p1"], rec["p2"]], 1)[None], dtype=torch.float64).cuda()
        pad = torch.tensor(np.array(rec["pad"])[None, None], dtype=torch.float64).cuda()
        par = torch.zeros((1, lay.P), dtype=torch.float64).cuda()
        ln = builder.sfc_boxes_batch(spec, pts, pad, par)
        row = par[0].cpu().numpy()[lay.p_sfc:lay.p_sfc + d * d + 2 * d]
        # atan2 / cos / sin of the device library vs glibc: last-place differences
        assert np.abs(row[:d * d].reshape(d, d) - np.array(rec["R"]).T).max() <= 1e-14
        assert np.abs(row[d * d:d * d + d] - np.array(rec["lower"])).max() <= 1e-13
        assert np.abs(row[d * d + d:] - np.array(rec["upper"])).max() <= 1e-13
        assert abs(ln[0, 0].item() - rec["length"]) <= 1e-14


@pytest.mark.gpu
def test_cuda_builders_reproduce_the_c4_batch(native_lib):
    """The 3-D corridor batch built on the device (boxes + initial guess) equals the vectorised host generator."""
    import torch
    from trajectory_generator_b200 import builder, synthetic as syn
    b = syn.make("C4", 2048)
    L = b.layout
    pts = torch.from_numpy(b.raw["points"]).cuda()
    dims = b.raw["dims"]                                    # [B, 4, 3] full dimensions: subtract the segment lengths
    seglen = np.linalg.norm(b.raw["points"][:, :, 1:] - b.raw["points"][:, :, :-1], 2, 1)
    pad = dims.copy(); pad[:, :, 0] -= seglen
    par = torch.from_numpy(b.par.copy()).cuda()
    par[:, L.p_sfc:] = 0
    builder.sfc_boxes_batch(b.spec, pts, torch.from_numpy(pad).cuda(), par)
    assert np.abs(par.cpu().numpy() - b.par).max() <= 1e-12
    x0 = builder.initial_guess_batch(b.spec, pts).cpu().numpy()
    assert np.abs(x0 - b.x0).max() <= 1e-12


@pytest.mark.gpu
def test_cuda_intervals_per_corridor_match_reference_fixtures(native_lib):
    """shape from geometry (DS/safe_flight_corridor.py:78-88) on the device, incl. the round-half-to-even ratios"""
    import torch
    from trajectory_generator_b200 import builder
    G = _golden()
    groups = {}
    for rec in G["intervals"]:
        pts = np.array(rec["points"])
        groups.setdefault((pts.shape, rec["min"]), []).append(rec)
    for (shape, mn), recs in groups.items():
        pts = torch.from_numpy(np.stack([np.array(r["points"]) for r in recs])).cuda()
        ipc, key = builder.sfc_intervals_batch(pts, mn)
        ipc = ipc.cpu().numpy(); key = key.cpu().numpy()
        for i, r in enumerate(recs):
            assert ipc[i].tolist() == r["ipc"], (shape, mn, i)
        for i in range(len(recs)):
            for j in range(len(recs)):
                assert (key[i] == key[j]) == (recs[i]["ipc"] == recs[j]["ipc"])


@pytest.mark.gpu
def test_corridor_problems_built_from_raw_geometry(native_lib):
    """f2 -> f3: raw corridor polylines in, shapes chosen on the device, one solve call for all shapes; every problem's
    answer equals the one-shape-at-a-time answer of the array-level API on the same rows, and the drop-in class
    (which lets SFC_Data choose the intervals on the host) lands on the same control points."""
    import torch
    from trajectory_generator_b200 import batch
    from trajectory_generator_b200.batched import CorridorProblems
    import helpers as h
    rng = np.random.default_rng(12)
    B, ncorr = 96, 3
    pts = np.zeros((B, 3, ncorr + 1))
    pts[:, :, 0] = rng.uniform(-5, 5, (B, 3))
    direction = rng.normal(size=(B, 3)); direction /= np.linalg.norm(direction, 2, 1)[:, None]
    for i in range(1, ncorr + 1):
        turn = rng.normal(size=(B, 3)) * 0.35
        direction = direction + turn; direction /= np.linalg.norm(direction, 2, 1)[:, None]
        pts[:, :, i] = pts[:, :, i - 1] + direction * rng.uniform(5, 11.5, B)[:, None]
    pads = np.stack([rng.uniform(2, 3, (B, ncorr)), rng.uniform(2, 3, (B, ncorr)), rng.uniform(2, 4, (B, ncorr))], 2)
    v0 = pts[:, :, 1] - pts[:, :, 0]; v0 /= np.linalg.norm(v0, 2, 1)[:, None]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    cp = CorridorProblems(3, corridor_points=t(pts), corridor_pads=t(pads), start_velocity=t(v0), end_zero_velocity=True,
                          max_velocity=5.0, max_acceleration=0.3, objective_function_type="minimal_velocity_path")
    shapes = cp.shapes()
    assert len(shapes) >= 3 and sum(c for _, c in shapes) == B
    out = cp.solve()
    assert (out["status"] == 0).double().mean().item() > 0.9
    seen = torch.zeros(B, dtype=torch.int32, device="cuda")
    for idx, prob, x in out["buckets"]:
        seen[idx] += 1
        one = batch.solve(prob.spec, prob.par, prob.x0.clone(), jacobian="fd")
        assert torch.equal(one["x"], x) and torch.equal(one["status"], out["status"][idx])
    assert (seen == 1).all()
    # the drop-in route on a few problems: containers -> SFC_Data chooses the same intervals -> same solution
    from trajectory_generator_b200.trajectory_generator import TrajectoryGenerator
    ns = h.product_namespace()
    col = lambda v: np.asarray(v, dtype=float).reshape(-1, 1)
    gen = TrajectoryGenerator(3)
    for idx, prob, x in out["buckets"][:3]:
        b = int(idx[0].item())
        sfcs = []
        for k in range(ncorr):
            R, T, Ln = ns["get3DRotationAndTranslationFromPoints"](col(pts[b, :, k]), col(pts[b, :, k + 1]))
            dims = pads[b, k].copy(); dims[0] += Ln
            sfcs.append(ns["SFC"](col(dims), T, R))
        sfc = ns["SFC_Data"](tuple(sfcs), pts[b], 1)
        assert [int(v) for v in sfc.get_intervals_per_corridor()] == cp.ipc[b].cpu().tolist()
        wd = ns["WaypointData"]((ns["Waypoint"](location=col(pts[b, :, 0]), velocity=col(v0[b])),
                                 ns["Waypoint"](location=col(pts[b, :, -1]), velocity=col([0, 0, 0]))))
        cc = ns["ConstraintsContainer"](wd, ns["DerivativeBounds"](5.0, 0.3), None, sfc, None)
        cps, scale, viol = gen.generate_trajectory(cc, "minimal_velocity_path")
        mine = x[0, :prob.d * prob.N].reshape(prob.d, prob.N).cpu().numpy()
        assert cps.shape == mine.shape
        if gen.last_result.status == 0 and int(out["status"][b].item()) == 0:
            assert np.abs(cps - mine).max() <= 1e-4          # inputs agree to ~1e-16; forward differences amplify that
