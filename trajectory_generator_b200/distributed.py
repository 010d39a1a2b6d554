"""Multi-GPU plumbing: independent problems are sharded across ranks (one process per GPU), the solve itself needs
no communication, and the packed result rows ``[x (n) | status | nit | is_violation | f]`` are collected with ONE
all-gather (NCCL over NVLink on the GPU box; the same code runs over gloo on CPU tensors in the tests)."""
import numpy as np


def shard_bounds(total, rank, world):
    """Contiguous slice [lo, hi) of rank `rank`; the first total % world ranks get one extra."""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_result_rows(torch, x, status, nit, violation, f):
    n = x.shape[1]
    rows = torch.empty((x.shape[0], n + 4), dtype=torch.float64, device=x.device)
    rows[:, :n] = x
    rows[:, n] = status
    rows[:, n + 1] = nit
    rows[:, n + 2] = violation
    rows[:, n + 3] = f
    return rows


def unpack_result_rows(rows, n):
    rows = np.asarray(rows)
    return dict(x=rows[:, :n], status=rows[:, n].astype(np.int32), nit=rows[:, n + 1].astype(np.int32),
                violation=rows[:, n + 2].astype(np.int32), f=rows[:, n + 3])


def all_gather_rows(rows, total=None, group=None):
    """Every rank receives the result rows of all problems in global order.  Shards may differ by one row
    (uneven split): they are padded to the largest shard for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return rows
    world = dist.get_world_size(group)
    if total is None:
        count = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
        dist.all_reduce(count, group=group)
        total = int(count.item())
    width = rows.shape[1]
    biggest = shard_bounds(total, 0, world)[1]
    padded = rows
    if rows.shape[0] < biggest:
        padded = torch.zeros((biggest, width), dtype=rows.dtype, device=rows.device)
        padded[:rows.shape[0]] = rows
    out = torch.empty((world * biggest, width), dtype=rows.dtype, device=rows.device)
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(total, r, world)
        parts.append(out[r * biggest:r * biggest + (hi - lo)])
    return torch.cat(parts, 0)


def solve_sharded(spec, par, x0, maxiter=100, ftol=1e-6, jacobian="fd", group=None):
    """Solve a batch that is replicated on every rank's HOST (numpy par [B,P], x0 [B,n]): each rank solves its
    contiguous shard on its own GPU and all ranks return the full result (dict of numpy arrays).  `jacobian`
    defaults to "fd" like the drop-in class (the mode that follows the reference's iterates)."""
    import torch
    import torch.distributed as dist
    from . import batch
    from .problem import Layout
    lay = Layout(spec)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(len(par), rank, world)
    dev = torch.device("cuda", torch.cuda.current_device())
    p = torch.from_numpy(np.ascontiguousarray(par[lo:hi])).to(dev)
    x = torch.from_numpy(np.ascontiguousarray(x0[lo:hi])).to(dev)
    out = batch.solve(spec, p, x, maxiter, ftol, jacobian)
    rows = pack_result_rows(torch, out["x"], out["status"], out["nit"], out["violation"], out["f"])
    full = all_gather_rows(rows, total=len(par), group=group)
    return unpack_result_rows(full.cpu().numpy(), lay.n)
