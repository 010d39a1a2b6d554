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
