"""N > 1 host logic on CPU: world_size-2 gloo processes shard a batch, exchange result rows with the one
all-gather the GPU path uses, and every rank ends with all rows in global order (even and uneven splits)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers  # noqa: F401
from trajectory_generator_b200 import distributed as tgd


def test_shard_bounds_cover_the_batch():
    for total in (0, 1, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 8):
            edges = [tgd.shard_bounds(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, total, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = tgd.shard_bounds(total, rank, world)
    idx = torch.arange(lo, hi, dtype=torch.float64)
    x = idx[:, None] * 10 + torch.arange(n, dtype=torch.float64)[None, :]
    rows = tgd.pack_result_rows(torch, x, idx % 3, idx % 5, idx % 2, idx * 0.5)
    full = tgd.all_gather_rows(rows, total=total)
    out = tgd.unpack_result_rows(full.numpy(), n)
    ok = (full.shape == (total, n + 4) and np.array_equal(out["x"][:, 0], np.arange(total) * 10.0)
          and np.array_equal(out["status"], np.arange(total) % 3) and np.array_equal(out["nit"], np.arange(total) % 5)
          and np.allclose(out["f"], np.arange(total) * 0.5))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [64, 65])
def test_all_gather_rows_world2_gloo(total):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
