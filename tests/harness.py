"""ctypes view of tests/_build/libmp3b_cpu_harness.so (product host code + HD math compiled for the CPU)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GCDESC_DTYPE = np.dtype([
    ("bit_off", "<u8"), ("part_23_length", "<u2"), ("big_values", "<u2"), ("scalefac_compress", "<u2"),
    ("global_gain", "u1"), ("flags0", "u1"), ("table_select", "u1", (3,)), ("region_count", "u1", (3,)),
    ("subblock_gain", "u1", (3,)), ("scfsi", "u1"), ("n_long_sfb", "u1"), ("n_short_sfb", "u1"),
    ("sr_idx", "u1"), ("flags1", "u1"), ("pad", "u1", (2,))])
TILEDESC_DTYPE = np.dtype([
    ("gc_first", "<u4"), ("n_gr", "u1"), ("nch", "u1"), ("pad", "u1", (2,)), ("pcm_off", "<u8"),
    ("halo", "<i4", (2, 2))])
assert GCDESC_DTYPE.itemsize == 32 and TILEDESC_DTYPE.itemsize == 32

_lib = None


def lib():
    global _lib
    if _lib is None:
        from minimp3_b200.build import build_cpu_harness
        _lib = C.CDLL(build_cpu_harness())
        _lib.h_pow43.restype = C.c_float
        _lib.h_pow43.argtypes = [C.c_int]
        _lib.h_ldexp_q2.restype = C.c_float
        _lib.h_ldexp_q2.argtypes = [C.c_float, C.c_int]
        _lib.h_detect_buf.argtypes = [C.c_void_p, C.c_size_t]
        vp = C.c_void_p
        _lib.h_walk.argtypes = [vp, C.c_size_t, C.c_int, C.c_int, vp, C.c_int, vp, C.c_int, vp, vp, vp, C.c_int, vp, vp]
        _lib.h_entropy.argtypes = [vp, C.c_size_t, vp, C.c_int, vp, vp, vp, vp, vp]
        _lib.h_transform.argtypes = [vp, vp, vp, vp, vp, C.c_int, vp, vp, vp, vp]
        _lib.h_dct32.argtypes = [vp]
        assert _lib.h_sizeof_gcdesc() == 32 and _lib.h_sizeof_tiledesc() == 32
    return _lib


def walk(data: bytes, raw=False, allow=False):
    L = lib()
    buf = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, np.uint8)
    sc = np.zeros(16, np.int64)
    L.h_walk(buf.ctypes.data, C.c_size_t(len(data)), int(raw), int(allow), None, 0, None, 0, None, None, None, 0, None,
             sc.ctypes.data)
    n_gc, n_tiles, n_gr = int(sc[11]), int(sc[12]), int(sc[13])
    gcs = np.zeros(max(1, n_gc), GCDESC_DTYPE)
    tiles = np.zeros(max(1, n_tiles), TILEDESC_DTYPE)
    gr_gc = np.zeros(max(1, n_gr), np.uint32)
    gr_nch = np.zeros(max(1, n_gr), np.uint8)
    gr_reset = np.zeros(max(1, n_gr), np.uint8)
    md = np.zeros(len(data) + 16, np.uint8)
    L.h_walk(buf.ctypes.data, C.c_size_t(len(data)), int(raw), int(allow), gcs.ctypes.data, n_gc, tiles.ctypes.data,
             n_tiles, gr_gc.ctypes.data, gr_nch.ctypes.data, gr_reset.ctypes.data, n_gr, md.ctypes.data, sc.ctypes.data)
    keys = ["ret", "samples_decoded", "out_skip", "out_samples", "channels", "hz", "layer", "avg_bitrate_kbps",
            "main_data_bytes", "frames", "frames_decoded", "n_gc", "n_tiles", "n_granules", "n_istereo"]
    info = {k: int(sc[i]) for i, k in enumerate(keys)}
    return dict(info=info, gcs=gcs[:n_gc], tiles=tiles[:n_tiles], gr_gc=gr_gc[:n_gr], gr_nch=gr_nch[:n_gr],
                gr_reset=gr_reset[:n_gr], md=md)


def entropy(md: np.ndarray, gcs: np.ndarray):
    L = lib()
    n = len(gcs)
    gcs = np.ascontiguousarray(gcs)
    q = np.zeros((n, 576), np.int16)
    xr = np.zeros((n, 576), np.float32)
    scf = np.zeros((n, 40), np.float32)
    ist = np.zeros((n, 40), np.uint8)
    ext = np.zeros(n, np.int32)
    L.h_entropy(md.ctypes.data, C.c_size_t(len(md)), gcs.ctypes.data, n, q.ctypes.data, xr.ctypes.data, scf.ctypes.data,
                ist.ctypes.data, ext.ctypes.data)
    return dict(q=q, xr=xr, scf=scf, ist=ist, extent=ext)


def transform(xr_stereo: np.ndarray, w, float_output=False):
    L = lib()
    n = len(w["gcs"])
    xr_stereo = np.ascontiguousarray(xr_stereo, dtype=np.float32)
    xr_aa = np.zeros((n, 576), np.float32)
    sub = np.zeros((n, 576), np.float32)
    total = int(sum(576*int(c) for c in w["gr_nch"]))
    pcm16 = np.zeros(max(1, total), np.int16)
    pcmf = np.zeros(max(1, total), np.float32)
    gcs = np.ascontiguousarray(w["gcs"])
    L.h_transform(xr_stereo.ctypes.data, gcs.ctypes.data, w["gr_gc"].ctypes.data, w["gr_nch"].ctypes.data,
                  w["gr_reset"].ctypes.data, len(w["gr_gc"]), xr_aa.ctypes.data, sub.ctypes.data, pcm16.ctypes.data,
                  pcmf.ctypes.data)
    return dict(xr_aa=xr_aa, sub=sub, pcm16=pcm16[:total], pcmf=pcmf[:total])
