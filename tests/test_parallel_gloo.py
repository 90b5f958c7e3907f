"""CPU, world_size 2 (gloo): the frame-sharding logic of
solex_ser_recon_en_b200.parallel -- per-rank partial sums / maxima all-reduce to
exactly the single-process mean frame, and per-rank reconstructed frame rows
gather into exactly the single-process disks.  The per-rank compute is done by
the oracle here (no GPU); on the B200 box the same functions run on device
tensors over NCCL."""
import os
import socket
import sys
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from helpers import case_stack, golden
        from oracle import shg_oracle as O
        from solex_ser_recon_en_b200 import parallel
        _, stack = case_stack('ser16_rot')
        g = golden('ser16_rot')
        n = stack.shape[0]
        k0, k1 = parallel.frame_range(n)
        assert parallel.world() == (rank, world)
        # pass 1: this rank's frames only
        s, m = O.raw_sum_max(stack, k0, k1)
        geom = types.SimpleNamespace(n_frames=n)
        part = types.SimpleNamespace(sum=torch.from_numpy(s.view(np.int64).ravel().copy()),
                                     max=torch.from_numpy(m.astype(np.int32).ravel().copy()), n=k1 - k0, geom=geom)
        tot_sum, tot_max, n_total = parallel.combine_stats(part)
        assert n_total == n
        mean_img, max_img = O.finalize_mean_max(tot_sum.numpy().view(np.uint64).reshape(s.shape),
                                                tot_max.numpy().reshape(m.shape).astype(np.uint16), n, False)
        assert np.array_equal(mean_img, g['mean_img']) and np.array_equal(max_img, g['max_img'])
        # pass 2: this rank's frame rows of every disk, gathered to rank 0
        shifts = [int(x) for x in g['shift']]
        disks = O.recon(stack, g['fit'], shifts, k0, k1)                       # list of (ih, k1-k0)
        local = torch.from_numpy(np.stack([d.T for d in disks]).copy())        # (S, n_local, ih) frame-major
        full = parallel.gather_rows(local, n, dst=0)
        if rank == 0:
            for i in range(len(shifts)):
                assert np.array_equal(full[i].numpy().T, g[f'disk{i}']), shifts[i]
        else:
            assert full is None
        # ownership maps and the small-object broadcast used for the ellipse geometry
        assert parallel.shift_owner(5, world, 'gather0') == [0] * 5
        own = parallel.shift_owner(101, world, 'by_shift')
        assert own[0] == 0 and own == sorted(own) and set(own) == set(range(world))
        geom = parallel.broadcast_object((1.5, [1, 2]) if rank == 0 else None, src=0)
        assert geom == (1.5, [1, 2])
        # the ellipse geometry travels as one tensor broadcast; a failed fit on rank 0 raises on EVERY rank
        parallel.device_barrier()
        want = ((101.25, 2047.5, 1638.125), 0.195, -0.0125, [3.0, 460.0, 3590.84, 3624.0])
        got = parallel.broadcast_geometry(want if rank == 0 else None, 0)
        assert got == want, got
        boom = ValueError('could not find any edges') if rank == 0 else None
        try:
            parallel.broadcast_geometry(None, 0, boom)
        except ValueError:
            assert rank == 0
        except Exception as e:
            assert rank != 0 and 'rank 0' in str(e)
        else:
            raise AssertionError('a failed fit must raise on every rank')
        parallel.device_barrier()
        # the same through the shared-memory mailbox the GPU runs use on one node (several rounds: sequence numbers)
        os.environ.update(LOCAL_WORLD_SIZE=str(world), SHG_GEOMETRY_MAILBOX='1')
        for rnd in range(3):
            w2 = ((want[0][0] + rnd, want[0][1], want[0][2]), want[1], want[2] * (rnd + 1), want[3])
            got = parallel.broadcast_geometry(w2 if rank == 0 else None, 0)
            assert got == w2, (rnd, got)
        try:
            parallel.broadcast_geometry(None, 0, ValueError('no edges') if rank == 0 else None)
        except ValueError:
            assert rank == 0
        except Exception as e:
            assert rank != 0 and 'rank 0' in str(e)
        else:
            raise AssertionError('a failed fit must raise on every rank')
        assert parallel._mailbox is not None
        dist.barrier()
        open(os.path.join(out_dir, 'ok%d' % rank), 'w').write('ok')
    finally:
        dist.destroy_process_group()


def test_two_ranks_reproduce_the_single_process_result(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(str(tmp_path))) == ['ok0', 'ok1']


def test_frame_ranges_tile_the_scan():
    sys.path.insert(0, ROOT)
    from solex_ser_recon_en_b200 import parallel
    for n in (1, 7, 100, 20000):
        for size in (1, 2, 3, 4, 8):
            r = [parallel.frame_range(n, g, size) for g in range(size)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(size - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_post_warp_pixel_ownership_tiles_the_image():
    """Exchange mode 'post_warp': a circularised pixel is produced by the rank whose (logical) frames contain
    its left tap floor(x).  For every world size / flip / tilt: exactly one producer per pixel, and the taps
    it reads are its own physical frames plus at most HALO frames of a neighbour."""
    sys.path.insert(0, ROOT)
    from solex_ser_recon_en_b200 import geometry, parallel
    n_frames, ih = 997, 160
    for phi, ratio in ((0.0, 0.31), (0.05, 0.25), (-0.2, 1.7)):
        _, mat3, (oh, ow), _, _ = geometry.warp_plan((ih, n_frames), phi, ratio)
        r, c = np.meshgrid(np.arange(oh), np.arange(ow), indexing='ij')
        kf = np.floor(mat3[0, 0] * c + mat3[0, 1] * r + mat3[0, 2]).astype(np.int64)
        for size in (2, 3, 8):
            for flip in (False, True):
                owners = np.zeros(kf.shape, np.int32)
                for rank in range(size):
                    lo, hi = parallel.owned_logical_frames(n_frames, rank, size, flip)
                    mine = (kf >= lo) & (kf < hi)
                    owners += mine
                    k0, k1 = parallel.frame_range(n_frames, rank, size)
                    taps = np.concatenate([kf[mine], kf[mine] + 1])
                    taps = taps[(taps >= 0) & (taps < n_frames)]            # outside the image: constant fill
                    phys = n_frames - 1 - taps if flip else taps
                    h = parallel.halo_frames(n_frames, size)
                    assert phys.size == 0 or (phys.min() >= k0 - h and phys.max() <= k1 - 1 + h), (size, flip, rank)
                assert np.all(owners == 1), (phi, size, flip)
