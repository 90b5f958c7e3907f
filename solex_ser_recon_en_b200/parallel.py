"""Frame-range sharding across the GPUs of one box.

One process per GPU (torchrun).  Rank g owns frames [g*N/G, (g+1)*N/G) of the
scan: it ingests and sums only those, the integer sum / max frames are combined
with an NCCL all-reduce over NVLink (exact: integer addition and max are
order-independent, so every world size gives identical bits), every rank runs
the tiny detection + fit on the combined frame, reconstructs its own frame rows
of every disk image, and the rows are gathered to rank 0.

With no process group (the normal single-GPU case) every function here is the
identity, so the drop-in modules call them unconditionally.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .engine import get_engine


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def frame_range(n_frames: int, rank: int | None = None, size: int | None = None):
    """Frames [k0, k1) owned by a rank: contiguous, sizes differ by at most one."""
    r, s = world()
    rank = r if rank is None else rank
    size = s if size is None else size
    return n_frames * rank // size, n_frames * (rank + 1) // size


def combine_stats(stack):
    """(sum, max, n_total) of the whole scan from every rank's partial sum / max."""
    _, size = world()
    if size == 1:
        return stack.sum, stack.max, stack.n
    total_sum = stack.sum.clone()
    total_max = stack.max.clone()
    dist.all_reduce(total_sum, op=dist.ReduceOp.SUM)          # int64 bit pattern of exact uint64 sums (< 2^63)
    dist.all_reduce(total_max, op=dist.ReduceOp.MAX)
    return total_sum, total_max, stack.geom.n_frames


def gather_rows(local_disk, n_frames: int, dst: int = 0):
    """Assemble the (n_shifts, N, ih) disk on rank `dst` from each rank's
    (n_shifts, n_local, ih) block of frame rows.  Returns the full tensor on
    `dst`, None elsewhere."""
    rank, size = world()
    if size == 1:
        return local_disk
    n_shifts, _, ih = local_disk.shape
    if rank == dst:
        full = torch.empty((n_shifts, n_frames, ih), dtype=local_disk.dtype, device=local_disk.device)
        k0, k1 = frame_range(n_frames, dst, size)
        full[:, k0:k1].copy_(local_disk)
        bufs = {}
        ops = []
        for src in range(size):
            if src == dst:
                continue
            a, b = frame_range(n_frames, src, size)
            bufs[src] = torch.empty((n_shifts, b - a, ih), dtype=local_disk.dtype, device=local_disk.device)
            ops.append(dist.P2POp(dist.irecv, bufs[src].view(torch.int16), src))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for src, buf in bufs.items():
            a, b = frame_range(n_frames, src, size)
            full[:, a:b].copy_(buf)
        return full
    for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, local_disk.contiguous().view(torch.int16), dst)]):
        w.wait()
    return None


def reconstruct(stack, fit: np.ndarray, shifts):
    """Disk images for the whole scan on rank 0 (every rank when single-GPU)."""
    eng = get_engine()
    local = eng.recon(stack, fit, shifts)
    rank, size = world()
    if size == 1:
        return local
    full = gather_rows(local, stack.geom.n_frames)
    return full if rank == 0 else local
