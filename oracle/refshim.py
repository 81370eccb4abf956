"""ctypes view of oracle/_ref/libminimp3_ref_*.so -- TEST INFRASTRUCTURE ONLY.

The shared objects are the UNMODIFIED reference (/root/reference/minimp3.h,
minimp3_ex.h) compiled behind oracle/ref_shim.c by oracle/Makefile.  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module; the product (minimp3_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE_SRC = os.environ.get("MINIMP3_REFERENCE", "/root/reference")


def build(force=False):
    """Build the reference shim if the reference sources are present; no-op otherwise."""
    want = [os.path.join(REF_DIR, n) for n in ("libminimp3_ref_s16.so", "libminimp3_ref_f32.so")]
    if not force and all(os.path.exists(w) for w in want):
        return True
    if not os.path.exists(os.path.join(REFERENCE_SRC, "minimp3.h")):
        return all(os.path.exists(w) for w in want)
    subprocess.check_call(["make", "-C", HERE, "ref", "REF=" + REFERENCE_SRC], stdout=subprocess.DEVNULL)
    return True


def available():
    return os.path.exists(os.path.join(REF_DIR, "libminimp3_ref_s16.so"))


class FrameRec(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("offset", "frame_bytes", "frame_offset", "samples", "channels", "hz", "layer", "bitrate_kbps")]


class GcRec(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("frame", "gr", "ch", "nch", "hdr", "bit_pos", "main_data_begin",
                 "part_23_length", "big_values", "scalefac_compress",
                 "global_gain", "block_type", "mixed_block_flag", "n_long_sfb", "n_short_sfb")] + [
        ("table_select", C.c_int*3), ("region_count", C.c_int*3), ("subblock_gain", C.c_int*3)] + [
        (n, C.c_int) for n in ("preflag", "scalefac_scale", "count1_table", "scfsi", "sfb_kind", "sr_idx", "pcm_offset")]


class Taps(C.Structure):
    _fields_ = [("cap", C.c_long), ("rec", C.POINTER(GcRec)), ("scf", C.c_void_p), ("ist_pos", C.c_void_p),
                ("q", C.c_void_p), ("xr", C.c_void_p), ("xr_stereo", C.c_void_p), ("xr_aa", C.c_void_p),
                ("sub", C.c_void_p)]


class Ref:
    """One flavour of the compiled reference: float_output=False -> int16 PCM (SSE2 path)."""

    def __init__(self, float_output=False):
        name = "libminimp3_ref_f32.so" if float_output else "libminimp3_ref_s16.so"
        path = os.path.join(REF_DIR, name)
        if not os.path.exists(path):
            build()
        self.lib = C.CDLL(path)
        self.float_output = float_output
        self.dtype = np.float32 if float_output else np.int16
        L = self.lib
        L.ref_decode_frames.restype = C.c_long
        L.ref_decode_frames.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_void_p, C.c_long,
                                        C.POINTER(C.c_long)]
        L.ref_load_buf.restype = C.c_int
        L.ref_load_buf.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                   C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_detect_buf.restype = C.c_int
        L.ref_detect_buf.argtypes = [C.c_void_p, C.c_size_t]
        L.ref_tap_stream.restype = C.c_long
        L.ref_tap_stream.argtypes = [C.c_void_p, C.c_long, C.POINTER(Taps), C.c_void_p, C.c_long, C.POINTER(C.c_long)]
        L.ref_synth_sequence.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.ref_imdct_sequence.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_dct_ii.argtypes = [C.c_void_p, C.c_int]
        L.ref_const_pow43.restype = C.c_float
        L.ref_const_pow43.argtypes = [C.c_int]
        L.ref_const_ldexp_q2.restype = C.c_float
        L.ref_const_ldexp_q2.argtypes = [C.c_float, C.c_int]
        L.ref_const_sfb_table.restype = C.c_int
        L.ref_const_sfb_table.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.ref_bench_decode.restype = C.c_long
        L.ref_bench_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]

    # ---- whole-stream decodes -------------------------------------------------
    def decode_frames(self, data: bytes):
        """Raw mp3dec_decode_frame loop.  Returns (pcm ndarray, list[FrameRec])."""
        buf = np.frombuffer(data, dtype=np.uint8)
        nf = C.c_long(0)
        total = self.lib.ref_decode_frames(buf.ctypes.data, len(buf), None, 0, None, 0, C.byref(nf))
        frames = (FrameRec*max(1, nf.value))()
        pcm = np.zeros(max(1, total), dtype=self.dtype)
        total2 = self.lib.ref_decode_frames(buf.ctypes.data, len(buf), pcm.ctypes.data, total, frames, nf.value,
                                            C.byref(nf))
        assert total2 == total
        return pcm[:total], list(frames)[:nf.value]

    def load_buf(self, data: bytes):
        """mp3dec_load_buf.  Returns (ret, pcm, dict(channels,hz,layer,avg_bitrate_kbps))."""
        buf = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, np.uint8)
        out = C.c_void_p()
        n = C.c_size_t(0)
        ch, hz, layer, br = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        ret = self.lib.ref_load_buf(buf.ctypes.data, len(data), C.byref(out), C.byref(n), C.byref(ch), C.byref(hz),
                                    C.byref(layer), C.byref(br))
        if out.value and n.value:
            arr = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_float if self.float_output else C.c_int16)),
                                        shape=(n.value,)).copy()
        else:
            arr = np.zeros(0, dtype=self.dtype)
        if out.value:
            self.lib.ref_free(out)
        return ret, arr, dict(channels=ch.value, hz=hz.value, layer=layer.value, avg_bitrate_kbps=br.value)

    def detect_buf(self, data: bytes):
        buf = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, np.uint8)
        return self.lib.ref_detect_buf(buf.ctypes.data, len(data))

    # ---- stage taps ------------------------------------------------------------
    def tap_stream(self, data: bytes, cap=None):
        """Stage taps for every Layer III granule-channel of a stream (see ref_shim.c)."""
        buf = np.frombuffer(data, dtype=np.uint8)
        tot = C.c_long(0)
        if cap is None:
            cap = self.lib.ref_tap_stream(buf.ctypes.data, len(buf), None, None, 0, C.byref(tot))
            if cap < 0:
                raise RuntimeError("ref_tap_stream self-check failed: %d" % cap)
        capn = max(1, cap)
        rec = (GcRec*capn)()
        arrs = dict(scf=np.zeros((capn, 40), np.float32), ist_pos=np.zeros((capn, 40), np.uint8),
                    q=np.zeros((capn, 576), np.int16), xr=np.zeros((capn, 576), np.float32),
                    xr_stereo=np.zeros((capn, 576), np.float32), xr_aa=np.zeros((capn, 576), np.float32),
                    sub=np.zeros((capn, 576), np.float32))
        t = Taps(capn, rec, *[arrs[k].ctypes.data for k in ("scf", "ist_pos", "q", "xr", "xr_stereo", "xr_aa", "sub")])
        self.lib.ref_tap_stream(buf.ctypes.data, len(buf), None, None, 0, C.byref(tot))
        pcm = np.zeros(max(1, tot.value), dtype=self.dtype)
        n = self.lib.ref_tap_stream(buf.ctypes.data, len(buf), C.byref(t), pcm.ctypes.data, len(pcm), C.byref(tot))
        if n < 0:
            raise RuntimeError("ref_tap_stream self-check failed: %d" % n)
        n = min(n, capn)
        out = {k: v[:n] for k, v in arrs.items()}
        out["rec"] = list(rec)[:n]
        out["pcm"] = pcm[:tot.value]
        return out

    def synth_sequence(self, sub: np.ndarray, nch: int):
        """sub: [n_gr, nch, 576] float32 subband samples -> pcm [n_gr*576*nch]."""
        sub = np.ascontiguousarray(sub, dtype=np.float32)
        n_gr = sub.shape[0]
        pcm = np.zeros(n_gr*576*nch, dtype=self.dtype)
        self.lib.ref_synth_sequence(sub.ctypes.data, n_gr, nch, pcm.ctypes.data)
        return pcm

    def imdct_sequence(self, xr_aa: np.ndarray, block_type, n_long_bands):
        xr_aa = np.ascontiguousarray(xr_aa, dtype=np.float32)
        n_gr = xr_aa.shape[0]
        bt = np.ascontiguousarray(block_type, dtype=np.int32)
        nl = np.ascontiguousarray(n_long_bands, dtype=np.int32)
        sub = np.zeros((n_gr, 576), np.float32)
        self.lib.ref_imdct_sequence(xr_aa.ctypes.data, n_gr, bt.ctypes.data, nl.ctypes.data, sub.ctypes.data)
        return sub

    def dct_ii(self, grbuf: np.ndarray, n=18):
        g = np.ascontiguousarray(grbuf, dtype=np.float32).copy()
        self.lib.ref_dct_ii(g.ctypes.data, n)
        return g

    # ---- CPU baseline ------------------------------------------------------------
    def bench_decode(self, streams, threads):
        """Decode `streams` (list of bytes) over `threads` pthreads; returns (samples, seconds)."""
        n = len(streams)
        keep = [np.frombuffer(s, dtype=np.uint8) for s in streams]
        ptrs = (C.c_void_p*n)(*[k.ctypes.data for k in keep])
        sizes = (C.c_long*n)(*[len(k) for k in keep])
        sec = C.c_double(0)
        total = self.lib.ref_bench_decode(ptrs, sizes, n, threads, C.byref(sec))
        return total, sec.value
