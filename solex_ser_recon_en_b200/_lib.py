"""ctypes binding of libshg.so (include/shg.h).  There is no fallback: if the
library has not been built, importing this module raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libshg.so')


class ShgError(Exception):
    """Raised when a libshg entry point returns non-zero (message from
    shg_last_error).  A plain Exception subclass so that the reference's
    ``try/except`` around a batch (SHG_MAIN.py:136-143) catches it."""


if not os.path.exists(LIB_PATH):
    raise ImportError(
        'libshg.so is missing (%s). Build it with `python -m solex_ser_recon_en_b200.build` '
        '(needs nvcc); there is no CPU fallback for the reconstruction path.' % LIB_PATH)

lib = C.CDLL(LIB_PATH)

vp, i32, i64, u64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double

_PROTOS = {
    'shg_last_error': (C.c_char_p, []),
    'shg_version': (i32, []),
    'shg_device_info': (i32, [i32, C.POINTER(i64)]),
    'shg_accumulate': (i32, [vp, i32, i64, i64, vp, vp, vp]),
    'shg_frame_sums': (i32, [vp, i32, i64, i64, vp, vp]),
    'shg_finalize_mean_max': (i32, [vp, vp, i64, i32, i32, i32, vp, vp, vp]),
    'shg_box_blur_u16': (i32, [vp, i32, i32, i32, i32, vp, vp, vp]),
    'shg_row_sums_u16': (i32, [vp, i32, i32, vp, vp]),
    'shg_row_argmin_u16': (i32, [vp, i32, i32, i32, i32, vp, vp]),
    'shg_polyfit3': (i32, [vp, vp, i32, i32, vp, vp, vp, vp]),
    'shg_sigma_mask': (i32, [vp, i32, dbl, vp, vp]),
    'shg_window_mask': (i32, [vp, i32, dbl, dbl, vp, vp]),
    'shg_fit_table': (i32, [vp, i32, vp, vp]),
    'shg_recon_workspace_bytes': (i64, [i32, i32]),
    'shg_recon': (i32, [vp, i32, i64, i32, i32, vp, vp, i32, vp, i64, vp, i64, i32, vp, i64, vp, C.POINTER(C.c_int), vp]),
    'shg_ipc_alloc': (i32, [i64, C.POINTER(vp), C.c_char_p]),
    'shg_ipc_free': (i32, [vp]),
    'shg_ipc_open': (i32, [C.c_char_p, C.POINTER(vp)]),
    'shg_ipc_close': (i32, [vp]),
    'shg_transpose_u16': (i32, [vp, i64, i64, vp, i32, vp]),
    'shg_minmax_u16': (i32, [vp, i64, i64, vp, i32, vp, vp]),
    'shg_checksum_u16': (i32, [vp, i64, vp, vp]),
    'shg_warp_rows': (i32, [vp, i64, vp, i32, i64, i32, i32, dbl, dbl, dbl, vp, vp, i64, i32, i32, vp]),
    'shg_warp_rows_window': (i32, [vp, i64, vp, i32, i64, i32, i32, dbl, dbl, dbl, vp, vp, i64, i32, i32, vp, i32, i32,
                                   vp, vp]),
    'shg_warp_rows_tma_ok': (i32, [vp, i64, i32, vp, i64, vp]),
    'shg_warp_rows_tma': (i32, [vp, i64, i32, i64, i64, vp, i32, i64, i32, i32, dbl, dbl, dbl, vp, vp, i64, i32, i32, vp,
                                i32, i32, vp, vp]),
    'shg_exchange_rows': (i32, [vp, i64, i32, i32, i32, dbl, dbl, dbl, i32, i32, vp, vp]),
    'shg_downscale4_sum': (i32, [vp, i64, i32, i32, vp, i32, i32, vp]),
    'shg_box_sum_u32': (i32, [vp, i32, i32, i32, i32, vp, vp, vp]),
    'shg_sum_u32': (i32, [vp, i64, vp, vp]),
    'shg_select_u32': (i32, [vp, i64, C.POINTER(i64), i32, C.POINTER(C.c_uint32), vp, vp]),
    'shg_blur_range': (i32, [vp, i64, dbl, dbl, vp, vp]),
    'shg_blur_hist': (i32, [vp, i64, dbl, dbl, C.POINTER(dbl), i32, vp, vp]),
    'shg_flood_smooth': (i32, [vp, i32, i32, dbl, dbl, C.POINTER(dbl), i32, dbl, vp, vp, vp]),
    'shg_sobel_mag': (i32, [vp, i32, i32, vp, vp, vp, vp]),
    'shg_nms_candidates': (i32, [vp, vp, vp, i32, i32, dbl, vp, C.c_uint32, vp, vp, vp]),
    'shg_limb_state_bytes': (i64, []),
    'shg_limb_front': (i32, [vp, i32, i32, i32, C.POINTER(i64), dbl, i32, vp, vp, vp, vp, C.POINTER(dbl), vp]),
    'shg_limb_canny': (i32, [vp, i32, i32, dbl, dbl, C.POINTER(dbl), i32, dbl, dbl, vp, vp, C.c_uint32, vp, vp,
                             C.c_uint32, vp, vp, vp, vp]),
    'shg_hull_vertices': (i32, [vp, i64, vp, C.POINTER(i64)]),
    'shg_conic_scatter': (i32, [vp, i64, vp]),
    'shg_label_points': (i32, [vp, i64, i64, vp, C.POINTER(C.c_int32)]),
    'shg_log_table': (i32, [vp, vp]),
    'shg_transv_workspace_bytes': (i64, [i32, i32, i32]),
    'shg_transv_row_stats': (i32, [vp, i32, i32, i32, i64, vp, vp, vp, i32, i32, vp, vp, i64, vp]),
    'shg_transv_gain': (i32, [vp, i32, i32, i32, vp, vp, i32, i32, vp, vp]),
    'shg_row_scale_u16': (i32, [vp, i32, i32, i32, i64, vp, vp, vp]),
    'shg_tile_hist_u16': (i32, [vp, i32, i32, i32, i32, vp, vp, vp]),
    'shg_clahe_tile_area': (i64, [i32, i32, i32, i32]),
    'shg_clahe_lut': (i32, [vp, i32, i64, dbl, vp, vp]),
    'shg_clahe_apply': (i32, [vp, i32, i32, i32, i32, vp, vp, vp, vp]),
    'shg_rescale_u16': (i32, [vp, i64, dbl, dbl, vp, vp]),
    'shg_ingest_create': (i32, [i32, i64, i32, i32, C.POINTER(vp)]),
    'shg_ingest_destroy': (i32, [vp]),
    'shg_ingest_file': (i32, [vp, C.c_char_p, i64, i64, i64, i64, i64, vp, i32, vp, vp, C.POINTER(dbl)]),
    'shg_ingest_memory': (i32, [vp, vp, i64, i64, i64, vp, i32, vp, vp, C.POINTER(dbl)]),
    'shg_host_alloc': (i32, [i64, C.POINTER(vp)]),
    'shg_host_free': (i32, [vp]),
    'shg_memcpy_async': (i32, [vp, vp, i64, i32, vp]),
    'shg_synth_fill': (i32, [vp, i32, i64, i64, i64, i32, i32, u64, vp]),
}

for _name, (_res, _args) in _PROTOS.items():
    _f = getattr(lib, _name)
    _f.restype = _res
    _f.argtypes = _args

EXPORTS = tuple(_PROTOS)


def check(rc: int) -> None:
    if rc != 0:
        raise ShgError(lib.shg_last_error().decode('utf-8', 'replace') or 'libshg error %d' % rc)


def call(name: str, *args) -> None:
    """Call an int-returning entry point and raise ShgError on failure."""
    check(getattr(lib, name)(*args))
