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


def shift_owner(n_shifts: int, size: int | None = None, mode: str = 'by_shift'):
    """Rank that ends up holding the complete image of each shift.
    'by_shift': contiguous blocks of the shift list (index 0, the ellipse-fit
    shift, always lands on rank 0), so circularisation / transversalium / output
    of the 101 images spread over the ranks like the reference's Pool workers;
    'gather0': every image on rank 0 (north_star's "rows are gathered to rank 0")."""
    _, s = world()
    size = s if size is None else size
    if mode == 'gather0' or size == 1:
        return [0] * n_shifts
    return [min(size - 1, j * size // n_shifts) for j in range(n_shifts)]


class _RawDeviceMemory:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {'shape': (int(nbytes),), 'typestr': '|u1', 'data': (int(ptr), False),
                                         'version': 2}


class RowExchange:
    """Peer-writable images for one scan geometry.  Every rank allocates the
    images it owns with cudaMalloc, publishes a CUDA IPC handle, and opens the
    other ranks' allocations: the reconstruction kernel of rank g then writes
    its frame rows [k0_g, k1_g) of EVERY shift straight into the owner's image
    through NVLink peer stores -- there is no separate gather / all-to-all pass
    and no staging copy."""

    def __init__(self, n_shifts, n_frames, ih, mode):
        import ctypes as C
        from ._lib import call
        self.eng = get_engine()
        self.rank, self.size = world()
        self.key = (n_shifts, n_frames, ih, mode)
        self.owner = shift_owner(n_shifts, self.size, mode)
        self.mine = [j for j in range(n_shifts) if self.owner[j] == self.rank]
        self.image_bytes = n_frames * ih * 2
        nbytes = max(1, len(self.mine)) * self.image_bytes
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        call('shg_ipc_alloc', nbytes, C.byref(ptr), handle)
        self.ptr = int(ptr.value)
        gathered = [None] * self.size
        dist.all_gather_object(gathered, bytes(handle.raw))
        self.bases = {}
        self.opened = []
        for r, h in enumerate(gathered):
            if r == self.rank:
                self.bases[r] = self.ptr
            else:
                p = C.c_void_p()
                call('shg_ipc_open', h, C.byref(p))
                self.bases[r] = int(p.value)
                self.opened.append(int(p.value))
        local_index = {}
        counts = [0] * self.size
        for j in range(n_shifts):
            local_index[j] = counts[self.owner[j]]
            counts[self.owner[j]] += 1
        self.ptrs = np.array([self.bases[self.owner[j]] + local_index[j] * self.image_bytes for j in range(n_shifts)],
                             dtype=np.uint64)
        raw = torch.as_tensor(_RawDeviceMemory(self.ptr, nbytes), device=self.eng.device)
        self.images = raw.view(torch.uint16).view(max(1, len(self.mine)), n_frames, ih)

    def close(self):
        from ._lib import lib
        for p in self.opened:
            lib.shg_ipc_close(p)
        self.opened = []
        if self.ptr:
            lib.shg_ipc_free(self.ptr)
            self.ptr = 0


_exchange = {}
EXCHANGE_MODE = 'by_shift'


def row_exchange(n_shifts, n_frames, ih, mode=None):
    mode = mode or EXCHANGE_MODE
    key = (n_shifts, n_frames, ih, mode)
    ex = _exchange.get('current')
    if ex is None or ex.key != key:
        if ex is not None:
            torch.cuda.synchronize()
            dist.barrier()
            ex.close()
        ex = RowExchange(n_shifts, n_frames, ih, mode)
        _exchange['current'] = ex
    return ex


def release_exchange():
    ex = _exchange.pop('current', None)
    if ex is not None:
        torch.cuda.synchronize()
        if dist.is_initialized():
            dist.barrier()
        ex.close()


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
            ops.append(dist.P2POp(dist.irecv, bufs[src].view(torch.uint8), src))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for src, buf in bufs.items():
            a, b = frame_range(n_frames, src, size)
            full[:, a:b].copy_(buf)
        return full
    for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, local_disk.contiguous().view(torch.uint8), dst)]):
        w.wait()
    return None


def reconstruct(stack, fit: np.ndarray, shifts, first_done=None):
    """Reconstruct this rank's frames at every shift.  Returns (images, mins, known):
    `images` has one frame-major (N, ih) device image per shift -- every image on a
    single GPU; with several ranks, the images this rank owns and None for the others.
    `mins` is an int32 device tensor with the minimum pixel of every image the
    reconstruction kernel tracked (`known[j]`; all ranks' rows included): the
    circularisation clips to it and does not need a pass over the images for it.

    `first_done(image0, ready_event)` (optional) is called once the image of shifts[0]
    -- the ellipse-fit shift -- is complete (on its owner, rank 0), while the
    other shifts are still being reconstructed: the caller starts the limb
    search there so that its host-side part hides under the big kernel."""
    eng = get_engine()
    rank, size = world()
    n_s = len(shifts)
    split = first_done is not None and n_s > 1
    mins = torch.full((n_s,), 65535, dtype=torch.int32, device=eng.device)
    known = [False] * n_s
    if size == 1:
        disk = eng.alloc_disk(n_s, stack.n, stack.geom.ih)
        if split:
            # one shift = two band rows per frame: the direct-load kernel (impl 1) beats a TMA tile that small
            eng.recon(stack, fit, shifts[:1], disk=disk[:1], k0_out=0, impl=1)
            ready = torch.cuda.Event()
            ready.record()
            # queue the big kernel BEFORE waking the helper thread: the helper's Python work would otherwise
            # hold the interpreter lock while this thread still has the launch to do
            eng.recon(stack, fit, shifts[1:], disk=disk[1:], k0_out=0, mins=mins[1:])
            known[1:] = [eng.recon_min_done] * (n_s - 1)
            first_done(disk[0], ready)
        else:
            eng.recon(stack, fit, shifts, disk=disk, k0_out=0, mins=mins)
            known = [eng.recon_min_done] * n_s
        return [disk[i] for i in range(n_s)], mins, known
    g = stack.geom
    ex = row_exchange(n_s, g.n_frames, g.ih)
    torch.cuda.synchronize()
    dist.barrier()                         # owners are done reading the previous scan's images
    if split:
        eng.recon(stack, fit, shifts[:1], out_ptrs=ex.ptrs[:1], k0_out=stack.k0, impl=1)
        torch.cuda.synchronize()
        dist.barrier()                     # every rank's rows of image 0 have landed on rank 0
        eng.recon(stack, fit, shifts[1:], out_ptrs=ex.ptrs[1:], k0_out=stack.k0, mins=mins[1:])
        known[1:] = [eng.recon_min_done] * (n_s - 1)
        if ex.owner[0] == rank:
            first_done(ex.images[0], None)             # image 0 is complete (barrier above)
    else:
        eng.recon(stack, fit, shifts, out_ptrs=ex.ptrs, k0_out=stack.k0, mins=mins)
        known = [eng.recon_min_done] * n_s
    # every rank tracked the minimum of its own frame rows (same kernel variant everywhere: same geometry)
    dist.all_reduce(mins, op=dist.ReduceOp.MIN)
    torch.cuda.synchronize()
    dist.barrier()                         # every rank's rows have landed
    out = [None] * n_s
    for n, j in enumerate(ex.mine):
        out[j] = ex.images[n]
    return out, mins, known


def broadcast_object(obj, src: int):
    """Small Python object from one rank to all (ellipse geometry)."""
    _, size = world()
    if size == 1:
        return obj
    box = [obj]
    dist.broadcast_object_list(box, src=src)
    return box[0]
