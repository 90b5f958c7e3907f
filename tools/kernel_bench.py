"""Per-kernel timing on a device-synthesised scan (development aid; bench.py is
the contract benchmark).  Prints achieved GB/s against the algorithmic bytes of
SURVEY.md 8(d).

    python tools/kernel_bench.py [--frames 20000] [--width 4096] [--height 512] [--shifts 50] [--reps 5]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from solex_ser_recon_en_b200.engine import ScanGeometry, get_engine   # noqa: E402


def timed(fn, reps, flush=None):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for _ in range(2):
        fn()
    for a, b in ev:
        if flush is not None:
            flush()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2], t[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=20000)
    ap.add_argument('--width', type=int, default=4096)
    ap.add_argument('--height', type=int, default=512)
    ap.add_argument('--bpp', type=int, default=2)
    ap.add_argument('--shifts', type=int, default=50)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--impl', type=int, default=0)
    ap.add_argument('--only', default='')
    a = ap.parse_args()
    eng = get_engine(0)
    geom = ScanGeometry(a.width, a.height, a.bpp, a.frames)
    st = eng.synth_stack(geom, seed=5)
    eng.sync()
    res = {}
    only = set(a.only.split(',')) if a.only else None

    def want(k):
        return only is None or k in only

    stack_bytes = a.frames * geom.frame_bytes
    if only is not None and 'ingest' in only:
        # file -> pinned ring -> HBM (+ overlapped pass 1) from a tmpfs SER file of min(frames, 2000) frames
        import time
        from solex_ser_recon_en_b200 import synth
        nf = min(a.frames, 2000)
        g2 = ScanGeometry(a.width, a.height, a.bpp, nf)
        path = '/dev/shm/shg_ingest_bench.SER'
        host = st.frames[:nf * geom.frame_bytes].cpu().numpy()
        with open(path, 'wb') as f:
            f.write(synth.ser_header(a.width, a.height, 8 * a.bpp, nf))
            f.write(host.tobytes())
        del host
        for threads, slot_mb, slots in ((1, 64, 4), (4, 64, 4), (8, 64, 4), (16, 64, 4), (8, 256, 4), (16, 256, 6)):
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                s2, stats = eng.ingest_file(path, g2, 178, n_threads=threads, slot_mb=slot_mb, n_slots=slots)
                best = min(best, time.perf_counter() - t0)
                del s2
            res['ingest_file t=%d slot=%dMB x%d' % (threads, slot_mb, slots)] = dict(
                ms=best * 1e3, GBps=nf * geom.frame_bytes / best / 1e9, read_wait_ms=stats[1] * 1e3)
        os.remove(path)
        print(json.dumps(dict(config=vars(a), results=res), indent=1))
        return
    if want('accumulate'):
        ms, best = timed(lambda: eng.accumulate(st), a.reps)
        res['accumulate'] = dict(ms=ms, best_ms=best, GBps=stack_bytes / ms / 1e6)
    eng.accumulate(st)
    mean_img, max_img = eng.finalize_mean_max(st.sum, st.max, st.n, geom)
    det = eng.detect_line(mean_img, max_img)
    fit = eng.fit_line(det, geom.ih)
    if want('detect'):
        ms, best = timed(lambda: (eng.finalize_mean_max(st.sum, st.max, st.n, geom), eng.detect_line(mean_img, max_img),
                                  eng.fit_line(det, geom.ih)), a.reps)
        res['finalize+detect+fit'] = dict(ms=ms, best_ms=best)
    shifts = list(dict.fromkeys([10, 0] + list(range(-a.shifts, a.shifts + 1))))
    disk = eng.alloc_disk(len(shifts), a.frames, geom.ih)
    nb = max(shifts) - min(shifts) + 2
    recon_bytes = a.frames * (geom.ih * nb * a.bpp + len(shifts) * geom.ih * 2)
    if want('recon'):
        mins = torch.full((len(shifts),), 65535, dtype=torch.int32, device=eng.device)   # tracked image minima
        ms, best = timed(lambda: eng.recon(st, fit['fit'], shifts, disk=disk, k0_out=0, impl=a.impl, mins=mins), a.reps)
        res['recon'] = dict(ms=ms, best_ms=best, GBps=recon_bytes / ms / 1e6, n_shifts=len(shifts), nb=nb)
    eng.recon(st, fit['fit'], shifts, disk=disk, k0_out=0)
    img_bytes = a.frames * geom.ih * 2
    if want('limb'):
        # the serial part of every multi-GPU step: limb search + ellipse fit of the first image, GPU otherwise idle
        import time
        from solex_ser_recon_en_b200 import ellipse_fit as E
        for label, chained in (('chained', True), ('stepwise', False)):
            ts = []
            for rep in range(a.reps + 2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                sums = eng.downscale4(disk[0], False)
                pts, raw = E.limb_points_device(eng, sums, chained=chained)
                t1 = time.perf_counter()
                E.two_step(pts * 4)
                t2 = time.perf_counter()
                ts.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3))
            ts = sorted(ts[2:])
            res['limb_' + label] = dict(search_ms=ts[len(ts) // 2][0], two_step_ms=ts[len(ts) // 2][1], points=len(pts))
    if want('minmax'):
        ms, best = timed(lambda: eng.minmax(disk[1]), a.reps)
        res['minmax'] = dict(ms=ms, best_ms=best, GBps=img_bytes / ms / 1e6)
    # a plausible correction: Y/X ratio from the synthetic ellipse, small tilt
    ratio = (0.40 * geom.ih) / (0.42 * a.frames)
    phi = 0.02
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from solex_ser_recon_en_b200 import geometry as G
    mat3, (oh, ow) = G.warp_plan((geom.ih, a.frames), phi, ratio)[1:3]
    lo, hi = eng.minmax(disk[1])
    circ = eng.empty((oh, ow), torch.uint16)
    if want('warp'):
        ms, best = timed(lambda: eng.warp(disk[1], False, mat3, (oh, ow), 300.0, lo, hi, out=circ), a.reps)
        res['warp'] = dict(ms=ms, best_ms=best, GBps=(img_bytes + oh * ow * 2) / ms / 1e6, out=(oh, ow))
    eng.warp(disk[1], False, mat3, (oh, ow), 300.0, lo, hi, out=circ)
    if only is not None and 'tbatch' in only:
        # row statistics of the whole batch alone (the launch of the real step), classic kernel beside it
        n_img = len(shifts)
        mm_all = eng.minmax_device(disk)
        circ_all = eng.empty((n_img, oh, ow), torch.uint16)
        eng.warp_batch(disk, None, False, mat3, (oh, ow), mm_all, out=circ_all)
        cy_, cx_, rad_ = oh / 2.0, ow / 2.0, 0.40 * geom.ih
        y1_, y2_, rows_, xa_, xb_ = eng.transversalium_chords((cx_, cy_, rad_), [0, 0, ow - 1, oh - 1])
        nbytes = n_img * 2.0 * float((xb_ - xa_).sum()) * 2
        outs = {}
        for label, env in (('reg', {}), ('reg_t256', {'SHG_TRANSV_T': '256'}),
                           ('classic', {'SHG_TRANSV_REG': '0'}), ('bitsliced', {'SHG_TRANSV_HIST': '0'})):
            os.environ.update(env)
            ms, best = timed(lambda: eng.transversalium_row_stats(circ_all, rows_, xa_, xb_, device=True), a.reps)
            outs[label] = eng.transversalium_row_stats(circ_all, rows_, xa_, xb_, device=True)
            for k in env:
                del os.environ[k]
            res['transv_stats_batch_' + label] = dict(ms=ms, best_ms=best, GBps=nbytes / ms / 1e6, rows=len(rows_),
                                                      images=n_img, max_len=int((xb_ - xa_).max()))
        for stop in (1, 2, 3, 4, 5, 6):                     # cumulative cost of the phases of the register-resident kernel
            os.environ['SHG_TRANSV_STOP'] = str(stop)
            ms, best = timed(lambda: eng.transversalium_row_stats(circ_all, rows_, xa_, xb_, device=True), a.reps)
            res['reg_stop_after_phase_%d' % stop] = round(ms, 3)
        del os.environ['SHG_TRANSV_STOP']
        res['identical_bits'] = dict(reg_vs_classic=bool(torch.equal(outs['reg'], outs['classic'])),
                                     reg_vs_bitsliced=bool(torch.equal(outs['reg'], outs['bitsliced'])))
        os.environ['SHG_TRANSV_REG'] = '1'
        eng.transversalium_row_stats(circ_all, rows_, xa_, xb_, device=True)
        res['rows_handed_back'] = eng._transv_work[:64].view(torch.int32).tolist()   # total, then per reason (transv.cu)
        print(json.dumps(dict(config=vars(a), results=res), indent=1))
        return
    if want('batch'):
        # the launches of the real step: every image of the scan in one call (what profiles/ must show)
        n_img = len(shifts)
        mm_all = eng.minmax_device(disk)
        circ_all = eng.empty((n_img, oh, ow), torch.uint16)
        ms, best = timed(lambda: eng.warp_batch(disk, None, False, mat3, (oh, ow), mm_all, out=circ_all), a.reps)
        res['warp_batch'] = dict(ms=ms, best_ms=best, GBps=n_img * (img_bytes + oh * ow * 2) / ms / 1e6, images=n_img,
                                 out=(oh, ow))
        os.environ['SHG_WARP_OLD'] = '1'
        ms, best = timed(lambda: eng.warp_batch(disk, None, False, mat3, (oh, ow), mm_all, out=circ_all), a.reps)
        res['warp_batch_direct_load_kernel'] = dict(ms=ms, best_ms=best, GBps=n_img * (img_bytes + oh * ow * 2) / ms / 1e6)
        del os.environ['SHG_WARP_OLD']
        cy_, cx_, rad_ = oh / 2.0, ow / 2.0, 0.40 * geom.ih
        y1_, y2_, rows_, xa_, xb_ = eng.transversalium_chords((cx_, cy_, rad_), [0, 0, ow - 1, oh - 1])
        ms, best = timed(lambda: eng.transversalium_row_stats(circ_all, rows_, xa_, xb_, device=True), a.reps)
        res['transv_stats_batch'] = dict(ms=ms, best_ms=best, GBps=n_img * 2.0 * float((xb_ - xa_).sum()) * 2 / ms / 1e6,
                                         rows=len(rows_), images=n_img)
        gains_all = torch.ones((n_img, oh), dtype=torch.float64, device=eng.device)
        det_all = eng.empty((n_img, oh, ow), torch.uint16)
        ms, best = timed(lambda: eng.row_scale(circ_all, gains_all, out=det_all), a.reps)
        res['row_scale_batch'] = dict(ms=ms, best_ms=best, GBps=n_img * 2 * oh * ow * 2 / ms / 1e6)
        del circ_all, det_all
    cy, cx, rad = oh / 2.0, ow / 2.0, 0.40 * geom.ih
    y1, y2, rows, xa, xb = eng.transversalium_chords((cx, cy, rad), [0, 0, ow - 1, oh - 1])
    if want('transv'):
        ms, best = timed(lambda: eng.transversalium_row_stats(circ, rows, xa, xb), a.reps)
        res['transv_stats'] = dict(ms=ms, best_ms=best, GBps=2.0 * float((xb - xa).sum()) * 2 / ms / 1e6,
                                   rows=len(rows), max_len=int((xb - xa).max()))
    gain = np.ones(oh)
    det_img = eng.empty((oh, ow), torch.uint16)
    if want('scale'):
        ms, best = timed(lambda: eng.row_scale(circ, gain, out=det_img), a.reps)
        res['row_scale'] = dict(ms=ms, best_ms=best, GBps=2 * oh * ow * 2 / ms / 1e6)
    if want('transpose'):
        ms, best = timed(lambda: eng.to_reference_layout(disk[1]), a.reps)
        res['transpose'] = dict(ms=ms, best_ms=best, GBps=2 * img_bytes / ms / 1e6)
    print(json.dumps(dict(config=vars(a), results=res), indent=1))


if __name__ == '__main__':
    main()
