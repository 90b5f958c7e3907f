"""Worker of tests/test_gpu_multi.py: one rank per GPU (torchrun), the drop-in
entry points on a golden scan with frames sharded across the ranks."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import torch
    import torch.distributed as dist
    name, out_dir, mode = sys.argv[1], sys.argv[2], sys.argv[3]
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
    from helpers import case_file, golden
    from oracle import ref_shim
    from oracle.make_golden import CASES
    from solex_ser_recon_en_b200 import Solex_recon, parallel
    parallel.EXCHANGE_MODE = mode
    g = golden(name)
    if rank == 0:
        case_file(name, out_dir)
    dist.barrier()
    path = os.path.join(out_dir, name + ('.avi' if CASES[name]['kind'] == 'avi' else '.SER'))
    case = CASES[name]
    got = {}

    def sink(basefich, image, cercle):
        got[int(basefich.rsplit('_shift=', 1)[1])] = np.asarray(image).copy()

    opt = ref_shim.default_options(shift=list(case['shift']), flip_x=case['flip_x'], ratio_fixe=case.get('ratio_fixe'),
                                   _result_sink=sink)
    opt['output_dir'] = out_dir
    for rep in range(2):                      # twice: the exchange buffers are reused
        got.clear()
        disk_list, bounds, hdr = Solex_recon.solex_read(path, dict(opt) if rep else opt)
        if rep == 0:
            shifts = [int(s) for s in opt['shift']]
            assert (int(bounds[0]), int(bounds[1])) == (int(g['y1']), int(g['y2']))
            if mode == 'post_warp':
                # every rank keeps its frame rows of every image; complete images exist only after the warp
                from solex_ser_recon_en_b200.device_image import PartialImage
                assert opt['_exchange'] == 'post_warp' and all(isinstance(d, PartialImage) for d in disk_list)
                k0, k1 = parallel.frame_range(int(g['disk0'].shape[1]), rank, world)
                for i, d in enumerate(disk_list):               # its own rows (between the halos) are the reference's
                    rows = d.tensor[d.halo:d.halo + (k1 - k0)].cpu().numpy().T
                    want = g[f'disk{i}'][:, ::-1] if case['flip_x'] else g[f'disk{i}']
                    assert np.array_equal(rows, want[:, k0:k1]), (rank, shifts[i])
                if rank == 0:
                    assert np.array_equal(np.asarray(disk_list[0].full), g['disk0'])
                req = [i for i, s in enumerate(shifts) if s in case['shift']]
                own = parallel.shift_owner(len(req), world, 'by_shift')
                mine = [shifts[i] for q, i in enumerate(req) if own[q] == rank]
            else:
                owner = parallel.shift_owner(len(shifts), world, mode)
                for i, d in enumerate(disk_list):
                    if owner[i] == rank:
                        assert d is not None and np.array_equal(np.asarray(d), g[f'disk{i}']), (rank, shifts[i])
                    else:
                        assert d is None
                mine = [s for i, s in enumerate(shifts) if owner[i] == rank and s in case['shift']]
            Solex_recon.solex_process(opt, disk_list, bounds, hdr)
            assert sorted(got) == sorted(mine), (rank, sorted(got), mine)
            for sh, det in got.items():
                d = np.abs(det.astype(np.int32) - g[f'det_{sh}'].astype(np.int32))
                assert d.max() <= 1 and np.mean(d != 0) < 1e-3, (rank, sh)
            np.testing.assert_allclose(float(opt['ratio_fixe']), float(g['ratio']), rtol=1e-9)
    parallel.release_exchange()
    open(os.path.join(out_dir, 'ok%d' % rank), 'w').write('ok')
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
