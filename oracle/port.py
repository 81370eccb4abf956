"""ctypes view of oracle/_build/libmp3_oracle.so (the plain-C restatement) -- TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libmp3_oracle.so")


class Taps(C.Structure):
    _fields_ = [("cap", C.c_long), ("q", C.c_void_p), ("scf", C.c_void_p), ("xr", C.c_void_p), ("sub", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        src = [os.path.join(HERE, n) for n in ("mp3_oracle.c", "mp3_oracle.h", "mp3_tables.h")]
        if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in src):
            subprocess.check_call(["make", "-C", HERE, "port"], stdout=subprocess.DEVNULL)
        _lib = C.CDLL(LIB)
        _lib.orc_decode_stream.restype = C.c_long
        _lib.orc_decode_stream.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_long, C.POINTER(Taps),
                                           C.POINTER(C.c_long)]
    return _lib


def decode_stream(data: bytes, taps=False, float_output=False):
    """Returns dict(pcm, n_gc[, q, scf, xr, sub])."""
    L = lib()
    buf = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, np.uint8)
    ngc = C.c_long(0)
    total = L.orc_decode_stream(buf.ctypes.data, len(data), None, None, 0, None, C.byref(ngc))
    if total < 0:
        return dict(ret=int(total), pcm=np.zeros(0, np.int16), n_gc=0)
    pcm16 = np.zeros(max(1, total), np.int16)
    pcmf = np.zeros(max(1, total), np.float32) if float_output else None
    out = dict(ret=0, n_gc=ngc.value)
    t = None
    if taps:
        n = max(1, ngc.value)
        arr = dict(q=np.zeros((n, 576), np.int16), scf=np.zeros((n, 40), np.float32), xr=np.zeros((n, 576), np.float32),
                   sub=np.zeros((n, 576), np.float32))
        t = Taps(n, *[arr[k].ctypes.data for k in ("q", "scf", "xr", "sub")])
        out.update({k: v[:ngc.value] for k, v in arr.items()})
    L.orc_decode_stream(buf.ctypes.data, len(data), pcm16.ctypes.data, pcmf.ctypes.data if float_output else None, total,
                        C.byref(t) if t is not None else None, C.byref(ngc))
    out["pcm"] = (pcmf if float_output else pcm16)[:total]
    return out
