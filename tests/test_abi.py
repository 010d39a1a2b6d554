"""The drop-in boundary: the CUDA library loads without a GPU, exports every symbol include/*.h declares
(the reference's 24 legacy symbols + the batched API), and every compute entry point FAILS LOUDLY when no device
is present -- there is no CPU fallback to fall into."""
import ctypes
import os
import re

import numpy as np
import pytest

import helpers
import legacy_abi

HEADER = os.path.join(helpers.ROOT, "include", "trajectory_generator_b200.h")


def _declared_symbols():
    text = open(HEADER).read()
    names = set(re.findall(r"\b(tg_[a-z_]+)\s*\(", text))
    legacy = set(re.findall(r"\b([A-Za-z_]+)_##D\s*\(", text))
    for base in legacy:
        names.add(base + "_2"); names.add(base + "_3")
    return names


def test_header_declares_the_24_reference_symbols():
    assert set(legacy_abi.SYMBOLS) <= _declared_symbols()
    assert len(legacy_abi.SYMBOLS) == 24


def test_library_exports_every_declared_symbol(native_lib):
    missing = [s for s in sorted(_declared_symbols()) if not hasattr(native_lib, s)]
    assert not missing, missing


def test_layout_mirror_matches_struct(native_lib):
    from trajectory_generator_b200 import _native, problem as pk
    assert native_lib.tg_spec_count() == pk.SP_COUNT
    spec = np.zeros(pk.SP_COUNT, dtype=np.int32)
    spec[pk.SP_DIM], spec[pk.SP_NCP], spec[pk.SP_NOBST], spec[pk.SP_START_VEL] = 2, 8, 3, 1
    lay = pk.Layout(spec)
    assert (lay.d, lay.N, lay.n, lay.n_obs, lay.meq) == (2, 8, 17, 3, 6)
    assert lay.P == 2 + 2 + 2 + 2 * 3 + 3
    assert len(_native.LAYOUT_FIELDS) == native_lib.tg_layout(None, None, 0)


def _no_gpu():
    import torch
    return not torch.cuda.is_available()


@pytest.mark.skipif(not _no_gpu(), reason="checks the no-device behaviour")
def test_compute_calls_fail_loudly_without_a_device(native_lib):
    from trajectory_generator_b200 import batch, problem as pk
    spec = np.zeros(pk.SP_COUNT, dtype=np.int32)
    spec[pk.SP_DIM], spec[pk.SP_NCP], spec[pk.SP_START_VEL], spec[pk.SP_END_VEL] = 2, 8, 1, 1
    lay = pk.Layout(spec)
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU path|CUDA"):
        batch.solve_host(spec, np.zeros((1, lay.P)), np.ones((1, lay.n)))
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU path|CUDA"):
        batch.evaluate_host(spec, np.zeros((1, lay.P)), np.ones((1, lay.n)))
    # legacy symbols have no status channel (reference ABI): they answer NaN, never a CPU-computed number
    k = dict(fn="get_spline_curvature_bound", D=2, N=4, pts=[0, 1, 2, 3, 0, 1, 0, 1])
    assert np.isnan(legacy_abi.call(native_lib, k)[0])


def test_product_never_imports_the_oracle():
    pkg = os.path.join(helpers.ROOT, "trajectory_generator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "tg_oracle" not in text and "hostsim" not in text.replace("tests/hostsim", ""), os.path.join(dirpath, f)
