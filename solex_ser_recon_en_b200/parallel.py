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

import math
import os

import numpy as np
import torch
import torch.distributed as dist

from ._lib import ShgError
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


_flag = {}


def device_barrier():
    """Rendezvous of the ranks' CURRENT STREAMS without blocking any host: a one-element NCCL all-reduce.
    Kernels queued behind it on any rank start only after every rank's stream has reached it, so peer
    stores issued before it (the row exchange inside the reconstruction kernel, the warp's stores into the
    owners' images) are complete and visible to whatever is queued after it.  It replaces the
    torch.cuda.synchronize() + dist.barrier() pairs of the first version of this module, each of which
    drained the GPU and left it idle for a host round trip."""
    _, size = world()
    if size == 1:
        return
    dev = _comm_device()
    t = _flag.get(dev)
    if t is None:
        t = _flag[dev] = torch.zeros(1, dtype=torch.int32, device=dev)
    if dev.type == 'cuda':
        with get_engine().stage('rendezvous'):
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)


def _comm_device():
    """Where collective payloads live: this rank's GPU under NCCL, host memory under gloo (CPU tests)."""
    if dist.get_backend() == 'gloo':
        return torch.device('cpu')
    return get_engine().device


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
        """Collective: every rank unmaps its peers' images, and only when ALL have done so does each rank
        free its own allocation (freeing memory a peer still has mapped, or closing a mapping whose memory
        the exporter already freed, leaves a CUDA error behind that the next checked call would report)."""
        from ._lib import call
        for p in self.opened:
            call('shg_ipc_close', p)
        self.opened = []
        if dist.is_initialized():
            dist.barrier()
        if self.ptr:
            call('shg_ipc_free', self.ptr)
            self.ptr = 0


_exchange = {}
# How the ranks exchange what they reconstruct (SHG_EXCHANGE overrides):
#  'by_shift'  every rank stores its frame rows of image j straight into the rank that owns shift j
#              (contiguous blocks of the shift list): complete disk images on their owners, 16.5 GB x (G-1)/G
#              over NVLink at config 5;
#  'gather0'   the same with every image on rank 0 (north_star's wording);
#  'post_warp' every rank keeps its frame rows of ALL images, circularises its own frame range (the warp is a
#              per-row resample along the frame axis) and stores the resulting column block of the 4-5 x smaller
#              circularised image into the owner: 3.6 GB x (G-1)/G over NVLink.  Only the ellipse-fit image is
#              gathered (rank 0 fits it).  Used by solex_read + solex_process when nothing needs the pixels of a
#              complete disk image on one GPU (no -f FITS of the raw disks, no display, no diagnostic plots,
#              no fixed ratio / tilt); otherwise 'by_shift' is used.
EXCHANGE_MODE = 'by_shift'


def exchange_mode():
    mode = os.environ.get('SHG_EXCHANGE', EXCHANGE_MODE)
    if mode not in ('by_shift', 'gather0', 'post_warp'):
        raise ShgError('SHG_EXCHANGE must be by_shift, gather0 or post_warp, not %r' % mode)
    return mode


def row_exchange(n_shifts, n_frames, ih, mode=None, slot='current'):
    mode = mode or (EXCHANGE_MODE if EXCHANGE_MODE != 'post_warp' else 'by_shift')
    key = (n_shifts, n_frames, ih, mode)
    ex = _exchange.get(slot)
    if ex is None or ex.key != key:
        if ex is not None:
            torch.cuda.synchronize()
            dist.barrier()
            ex.close()
        ex = RowExchange(n_shifts, n_frames, ih, mode)
        _exchange[slot] = ex
    return ex


def release_exchange():
    for slot in list(_exchange):
        ex = _exchange.pop(slot)
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
    device_barrier()                       # owners are done reading the previous scan's images
    if split:
        eng.recon(stack, fit, shifts[:1], out_ptrs=ex.ptrs[:1], k0_out=stack.k0, impl=1)
        device_barrier()                   # every rank's rows of image 0 have landed on rank 0
        ready = torch.cuda.Event()
        ready.record()
        eng.recon(stack, fit, shifts[1:], out_ptrs=ex.ptrs[1:], k0_out=stack.k0, mins=mins[1:], impl=_beside_fit())
        known[1:] = [eng.recon_min_done] * (n_s - 1)
        if ex.owner[0] == rank:
            first_done(ex.images[0], ready)            # image 0 is complete once `ready` has passed
    else:
        eng.recon(stack, fit, shifts, out_ptrs=ex.ptrs, k0_out=stack.k0, mins=mins)
        known = [eng.recon_min_done] * n_s
    # every rank tracked the minimum of its own frame rows (same kernel variant everywhere: same geometry);
    # the all-reduce doubles as the rendezvous after which every rank's rows have landed
    dist.all_reduce(mins, op=dist.ReduceOp.MIN)
    out = [None] * n_s
    for n, j in enumerate(ex.mine):
        out[j] = ex.images[n]
    return out, mins, known


def _beside_fit() -> int:
    """shg_recon's impl word for the big reconstruction on several GPUs: the limb search of the ellipse fit runs
    beside it on rank 0 and IS the critical path there (every rank waits for the geometry), so the persistent
    reconstruction kernel leaves it a quarter of the SMs (SHG_RECON_SM_CAP overrides; 0 = no cap).  The same cap
    on every rank keeps the kernel variant -- and with it the tracked minima -- identical everywhere."""
    cap = int(os.environ.get('SHG_RECON_SM_CAP', '112'))
    return (cap & 0xff) << 8


# ----------------------------------------------------------------------------
# exchange mode 'post_warp'
HALO = 2          # frames of each neighbour kept next to a rank's own: the warp of a pixel whose left tap
#                   floor(x) is this rank's frame reads frame floor(x) + 1 at most (see shg_warp_rows_window)


def halo_frames(n_frames: int, size: int) -> int:
    return max(1, min(HALO, n_frames // size))


def _halo_exchange(local, h: int, n_local: int):
    """local: (S, h + n_local + h, ih).  Fill the margins with the neighbours' border frames (NCCL P2P)."""
    rank, size = world()
    ops, recv = [], {}
    if rank > 0:
        send = local[:, h:2 * h].contiguous()
        recv['l'] = torch.empty_like(send)
        ops += [dist.P2POp(dist.isend, send.view(torch.uint8), rank - 1),
                dist.P2POp(dist.irecv, recv['l'].view(torch.uint8), rank - 1)]
    if rank < size - 1:
        send2 = local[:, n_local:n_local + h].contiguous()
        recv['r'] = torch.empty_like(send2)
        ops += [dist.P2POp(dist.isend, send2.view(torch.uint8), rank + 1),
                dist.P2POp(dist.irecv, recv['r'].view(torch.uint8), rank + 1)]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    if 'l' in recv:
        local[:, :h].copy_(recv['l'])
    if 'r' in recv:
        local[:, h + n_local:].copy_(recv['r'])


def reconstruct_partial(stack, fit: np.ndarray, shifts, first_done=None):
    """Exchange mode 'post_warp': reconstruct this rank's frames of EVERY shift into a local buffer with a
    halo of neighbour frames on each side.  Returns a list of device_image.PartialImage (one per shift, on
    every rank).  The image of shifts[0] (the ellipse-fit shift) is ALSO gathered on rank 0 -- a single image,
    by peer stores -- where `first_done(image0, None)` starts the limb search; it is attached as .full there."""
    from .device_image import DeviceImage, PartialImage
    eng = get_engine()
    rank, size = world()
    g = stack.geom
    n_s, n_local, n_frames, ih = len(shifts), stack.n, g.n_frames, g.ih
    h = halo_frames(n_frames, size)
    local = eng.empty((n_s, n_local + 2 * h, ih), torch.uint16)
    mins = torch.full((n_s,), 65535, dtype=torch.int32, device=eng.device)
    split = first_done is not None and n_s > 1
    if split:
        ex0 = row_exchange(1, n_frames, ih, mode='gather0', slot='first')
        device_barrier()                   # rank 0 is done reading the previous scan's image
        eng.recon(stack, fit, shifts[:1], out_ptrs=ex0.ptrs[:1], k0_out=stack.k0, impl=1)
        device_barrier()                   # every rank's rows of image 0 have landed on rank 0
        ready = torch.cuda.Event()
        ready.record()
    eng.recon(stack, fit, shifts, disk=local, k0_out=h, mins=mins, impl=_beside_fit() if split else 0)
    if not eng.recon_min_done:             # kernel variant without minimum tracking: one pass over the local rows
        mins = eng.minmax_device(local[:, h:h + n_local])[:, 0].contiguous()
    if split and rank == 0:
        first_done(ex0.images[0], ready)
    _halo_exchange(local, h, n_local)
    # whole-image minimum and the two candidate [0][0] pixels (first / last frame, slit position 0): one all-reduce
    red = torch.zeros((3, n_s), dtype=torch.int32, device=eng.device)
    red[0] = -mins
    as_i16 = local.view(torch.int16)       # (uint16 -> int32 through int16 + mask: plain dtype conversions only)
    if stack.k0 == 0:
        red[1] = as_i16[:, h, 0].to(torch.int32) & 0xFFFF
    if stack.k0 + n_local == n_frames:
        red[2] = as_i16[:, h + n_local - 1, 0].to(torch.int32) & 0xFFFF
    dist.all_reduce(red, op=dist.ReduceOp.MAX)
    red[0] = -red[0]
    mins_all = red[0]
    parts = []
    for j in range(n_s):
        p = PartialImage(eng, local[j], stack.k0, stack.k0 + n_local, h, n_frames)
        p.min_ref = (mins_all, j)
        p.cval_ref = (red, j)
        parts.append(p)
    if split and rank == 0:
        parts[0].full = DeviceImage(eng, ex0.images[0], 'frames')
    return parts


INT_MIN, INT_MAX = -2 ** 31, 2 ** 31 - 1


def owned_logical_frames(n_frames: int, rank: int, size: int, flip: bool):
    """[lo, hi) of the frames rank `rank` holds, in the order the image shows them (a flipped image shows
    physical frame p at N-1-p), with INT_MIN / INT_MAX for the two ends of the image: the pixels left of
    frame 0 and right of frame N-1 (constant fill) belong to the first / last range.  The ranges of all
    ranks tile the integers, so every output pixel of the circularisation has exactly one producer."""
    k0, k1 = frame_range(n_frames, rank, size)
    lo, hi = (n_frames - k1, n_frames - k0) if flip else (k0, k1)
    return (INT_MIN if lo == 0 else lo), (INT_MAX if hi == n_frames else hi)


def circ_exchange(n_imgs: int, out_rows: int, out_cols: int):
    """Peer-writable circularised images (n_imgs of out_rows x out_cols), owners by position in the list."""
    ex = row_exchange(n_imgs, out_rows, out_cols, mode='by_shift', slot='circ')
    if getattr(ex, 'ptrs_dev', None) is None:
        ex.ptrs_dev = get_engine().upload(ex.ptrs.view(np.int64)).view(torch.int64)
    return ex


def broadcast_object(obj, src: int):
    """Small Python object from one rank to all."""
    _, size = world()
    if size == 1:
        return obj
    box = [obj]
    dist.broadcast_object_list(box, src=src)
    return box[0]


class _Mailbox:
    """A ring of eight 12-double slots in a shared mapping (a file in /dev/shm, unlinked once every rank has mapped
    it): slot = [sequence number, status + nine geometry values, spare].  The ranks of this package always share one
    box (frames are sharded over the GPUs of ONE node), so the ellipse geometry can travel through host memory: the
    writer fills slot seq % 8, then stores the sequence number; the readers -- which have nothing else to do until
    the geometry exists -- poll it.  ~2 us instead of the ~200 us of an NCCL broadcast bracketed by an upload and a
    blocking download.  (The writer cannot lap a reader by eight scans: every scan has collectives.)"""
    SLOTS = 8

    def __init__(self):
        import mmap
        import tempfile
        rank, size = world()
        path = None
        nbytes = self.SLOTS * 12 * 8
        if rank == 0:
            base = '/dev/shm' if os.path.isdir('/dev/shm') else tempfile.gettempdir()
            fd, path = tempfile.mkstemp(prefix='shg_geometry_', dir=base)
            os.write(fd, bytes(nbytes))
            os.close(fd)
        path = broadcast_object(path, 0)
        with open(path, 'r+b') as f:
            self.map = mmap.mmap(f.fileno(), nbytes)
        self.buf = np.frombuffer(self.map, dtype=np.float64).reshape(self.SLOTS, 12)
        dist.barrier()                                      # everybody has it mapped: the name can go
        if rank == 0:
            os.unlink(path)
        self.seq = 0

    def post(self, vals10):
        self.seq += 1
        slot = self.buf[self.seq % self.SLOTS]
        slot[1:11] = vals10
        slot[0] = float(self.seq)                           # x86 keeps the store order: payload first

    def take(self):
        import time
        self.seq += 1
        slot = self.buf[self.seq % self.SLOTS]
        want = float(self.seq)
        spins = 0
        while slot[0] != want:
            spins += 1
            if spins > 200:
                time.sleep(2e-5)                            # (lets a worker thread of this process run)
        return np.array(slot[1:11])


_mailbox = None


def _single_node() -> bool:
    return int(os.environ.get('LOCAL_WORLD_SIZE', '0')) == int(os.environ.get('WORLD_SIZE', '-1'))


def broadcast_geometry(geom, src: int = 0, error: BaseException | None = None):
    """The ellipse geometry (circle (cx, cy, r), ratio, phi, borders [4]) from the rank that fitted it to
    all: nine doubles and a status word through a shared-memory mailbox (one node: the normal case) or in ONE
    NCCL broadcast of a device tensor (broadcast_object_list pickles and costs two broadcasts and two host
    synchronisations).  If the fit failed on `src` (`error`), every rank raises instead of waiting in a collective
    for a result that will not come."""
    global _mailbox
    rank, size = world()
    if size == 1:
        if error is not None:
            raise error
        return geom
    vals = np.zeros(10, dtype=np.float64)
    if rank == src:
        if error is None:
            circle, ratio, phi, borders = geom
            vals[:9] = [circle[0], circle[1], circle[2], ratio, phi] + [float(b) for b in borders]
            vals[9] = 1.0
    if src == 0 and _single_node() and (dist.get_backend() != 'gloo' or os.environ.get('SHG_GEOMETRY_MAILBOX')) \
            and not os.environ.get('SHG_GEOMETRY_NCCL'):
        if _mailbox is None:
            _mailbox = _Mailbox()
        if rank == 0:
            _mailbox.post(vals)
            v = vals
        else:
            v = _mailbox.take()
    else:
        dev = _comm_device()
        if rank == src:
            t = get_engine().upload(vals) if dev.type == 'cuda' else torch.from_numpy(vals)
        else:
            t = torch.empty((10,), dtype=torch.float64, device=dev)
        dist.broadcast(t, src=src)
        v = t.cpu().numpy()
    if v[9] != 1.0:
        if error is not None:
            raise error
        raise ShgError('the ellipse fit failed on rank %d (see its traceback)' % src)
    return (float(v[0]), float(v[1]), float(v[2])), float(v[3]), float(v[4]), [float(b) for b in v[5:9]]
