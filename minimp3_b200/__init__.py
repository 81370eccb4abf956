"""minimp3_b200 -- B200 (sm_100a) batched MP3 Layer III decoder behind the lieff/minimp3 API.

The product is the C-ABI shared library ``libminimp3_b200.so`` (include/minimp3_b200.h); this package
is the thin ctypes mirror used by the tests and bench.py.  Names and argument meaning follow the
reference (minimp3.h / minimp3_ex.h): ``mp3dec_init``, ``mp3dec_decode_frame``, ``mp3dec_load_buf``,
``mp3dec_detect_buf`` plus the batch entry ``mp3dec_decode_batch_cuda``.

There is no CPU decode path: if the CUDA library is missing or no device is usable, calls raise.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libminimp3_b200.so")

MP3D_E_PARAM, MP3D_E_MEMORY, MP3D_E_IOERROR, MP3D_E_USER, MP3D_E_DECODE = -1, -2, -3, -4, -5
MP3D_E_UNSUPPORTED, MP3D_E_CUDA = -6, -7
OUT_MALLOC, OUT_PINNED, OUT_DEVICE = 0, 1, 2
MINIMP3_MAX_SAMPLES_PER_FRAME = 1152*2


class FrameInfo(C.Structure):
    """mp3dec_frame_info_t (minimp3.h:13-16)"""
    _fields_ = [(n, C.c_int) for n in ("frame_bytes", "frame_offset", "channels", "hz", "layer", "bitrate_kbps")]


class Mp3Dec(C.Structure):
    """mp3dec_t (minimp3.h:18-23), 6668 bytes"""
    _fields_ = [("mdct_overlap", C.c_float*(2*9*32)), ("qmf_state", C.c_float*(15*2*32)), ("reserv", C.c_int),
                ("free_format_bytes", C.c_int), ("header", C.c_ubyte*4), ("reserv_buf", C.c_ubyte*511)]


class FileInfo(C.Structure):
    """mp3dec_file_info_t (minimp3_ex.h:38-43)"""
    _fields_ = [("buffer", C.c_void_p), ("samples", C.c_size_t), ("channels", C.c_int), ("hz", C.c_int),
                ("layer", C.c_int), ("avg_bitrate_kbps", C.c_int)]


class BatchStream(C.Structure):
    _fields_ = [("buf", C.c_void_p), ("size", C.c_size_t)]


class BatchOpts(C.Structure):
    _fields_ = [("device", C.c_int), ("float_output", C.c_int), ("raw_frames", C.c_int),
                ("allow_mono_stereo_transition", C.c_int), ("host_threads", C.c_int), ("output", C.c_int),
                ("cuda_stream", C.c_void_p), ("n_devices", C.c_int), ("devices", C.c_int*8), ("gpu_frame_scan", C.c_int), ("pinned_inputs", C.c_void_p), ("pinned_inputs_bytes", C.c_size_t)]


class BatchStats(C.Structure):
    _fields_ = [("input_bytes", C.c_uint64), ("samples", C.c_uint64), ("n_granule_channels", C.c_uint64),
                ("n_tiles", C.c_uint64), ("n_istereo_granules", C.c_uint64), ("n_frames", C.c_uint64),
                ("h2d_bytes", C.c_uint64),
                ("kernels_per_run", C.c_int), ("host_parse_ms", C.c_double), ("h2d_ms", C.c_double),
                ("kernels_prep", C.c_int)]


class BatchCallStats(C.Structure):
    _fields_ = [("total_ms", C.c_double), ("host_parse_ms", C.c_double), ("issue_ms", C.c_double),
                ("finish_ms", C.c_double), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("n_chunks", C.c_int), ("n_devices", C.c_int), ("scanned_streams", C.c_int)]


EXPORTS = ["mp3dec_init", "mp3dec_decode_frame", "mp3dec_decode_frame_f32", "mp3dec_f32_to_s16", "mp3dec_detect_buf",
           "mp3dec_load_buf", "mp3dec_load_buf_f32", "mp3dec_decode_batch_cuda", "mp3dec_batch_plan_create",
           "mp3dec_batch_plan_run", "mp3dec_batch_plan_run_stage", "mp3dec_batch_plan_sync", "mp3dec_batch_plan_fetch", "mp3dec_batch_plan_stats",
           "mp3dec_batch_plan_device_pcm", "mp3dec_batch_plan_destroy", "mp3dec_batch_plan_enable_taps",
           "mp3dec_batch_plan_read_tap", "mp3dec_b200_pow43_device", "mp3dec_batch_shard", "mp3dec_b200_version",
           "mp3dec_detect_cb", "mp3dec_load_cb", "mp3dec_iterate_buf", "mp3dec_iterate_cb", "mp3dec_ex_open_buf",
           "mp3dec_ex_open_cb", "mp3dec_ex_close", "mp3dec_ex_seek", "mp3dec_ex_read_frame", "mp3dec_ex_read",
           "mp3dec_detect", "mp3dec_load", "mp3dec_iterate", "mp3dec_ex_open", "mp3dec_batch_last_call_stats", "mp3dec_batch_release_outputs"]


class MapInfo(C.Structure):
    _fields_ = [("buffer", C.c_void_p), ("size", C.c_size_t)]


class IndexT(C.Structure):
    _fields_ = [("frames", C.c_void_p), ("num_frames", C.c_size_t), ("capacity", C.c_size_t)]


class Mp3DecEx(C.Structure):
    """mp3dec_ex_t (minimp3_ex.h:74-88), int16 flavour"""
    _fields_ = [("mp3d", Mp3Dec), ("file", MapInfo), ("io", C.c_void_p), ("index", IndexT),
                ("offset", C.c_uint64), ("samples", C.c_uint64), ("detected_samples", C.c_uint64),
                ("cur_sample", C.c_uint64), ("start_offset", C.c_uint64), ("end_offset", C.c_uint64),
                ("info", FrameInfo), ("buffer", C.c_int16*MINIMP3_MAX_SAMPLES_PER_FRAME),
                ("input_consumed", C.c_size_t), ("input_filled", C.c_size_t),
                ("is_file", C.c_int), ("flags", C.c_int), ("vbr_tag_found", C.c_int), ("indexes_built", C.c_int),
                ("free_format_bytes", C.c_int), ("buffer_samples", C.c_int), ("buffer_consumed", C.c_int),
                ("to_skip", C.c_int), ("start_delay", C.c_int), ("last_error", C.c_int)]

_lib = None


def load():
    """Load libminimp3_b200.so (built in-tree by minimp3_b200/build.py).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libminimp3_b200.so not built (run `python -m minimp3_b200.build`); "
                           "there is no CPU fallback for the decode path")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.mp3dec_init.argtypes = [C.POINTER(Mp3Dec)]
    L.mp3dec_init.restype = None
    for fn in (L.mp3dec_decode_frame, L.mp3dec_decode_frame_f32):
        fn.argtypes = [C.POINTER(Mp3Dec), vp, C.c_int, vp, C.POINTER(FrameInfo)]
        fn.restype = C.c_int
    L.mp3dec_f32_to_s16.argtypes = [vp, vp, C.c_int]
    L.mp3dec_f32_to_s16.restype = None
    L.mp3dec_detect_buf.argtypes = [vp, C.c_size_t]
    for fn in (L.mp3dec_load_buf, L.mp3dec_load_buf_f32):
        fn.argtypes = [C.POINTER(Mp3Dec), vp, C.c_size_t, C.POINTER(FileInfo), vp, vp]   # progress_cb: PROGRESS_CB or None
        fn.restype = C.c_int
    L.mp3dec_decode_batch_cuda.argtypes = [C.POINTER(BatchStream), C.c_int, C.POINTER(FileInfo), C.POINTER(C.c_int),
                                           C.POINTER(BatchOpts)]
    L.mp3dec_batch_plan_create.argtypes = [C.POINTER(BatchStream), C.c_int, C.POINTER(BatchOpts), C.POINTER(C.c_int)]
    L.mp3dec_batch_plan_create.restype = vp
    for name in ("mp3dec_batch_plan_run", "mp3dec_batch_plan_run_stage", "mp3dec_batch_plan_sync", "mp3dec_batch_plan_enable_taps"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = C.c_int
    L.mp3dec_batch_plan_run_stage.argtypes = [vp, C.c_int]
    L.mp3dec_batch_plan_run_stage.restype = C.c_int
    L.mp3dec_batch_plan_fetch.argtypes = [vp, C.POINTER(FileInfo), C.POINTER(C.c_int)]
    L.mp3dec_batch_plan_stats.argtypes = [vp, C.POINTER(BatchStats)]
    L.mp3dec_batch_plan_stats.restype = None
    L.mp3dec_batch_release_outputs.argtypes = [C.c_int]
    L.mp3dec_batch_release_outputs.restype = None
    L.mp3dec_batch_last_call_stats.argtypes = [C.POINTER(BatchCallStats)]
    L.mp3dec_batch_last_call_stats.restype = None
    L.mp3dec_batch_plan_device_pcm.argtypes = [vp]
    L.mp3dec_batch_plan_device_pcm.restype = vp
    L.mp3dec_batch_plan_destroy.argtypes = [vp]
    L.mp3dec_batch_plan_destroy.restype = None
    L.mp3dec_batch_plan_read_tap.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.mp3dec_b200_pow43_device.argtypes = [vp, C.c_int]
    L.mp3dec_b200_pow43_device.restype = C.c_int
    L.mp3dec_batch_shard.argtypes = [C.POINTER(C.c_size_t), C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.mp3dec_batch_shard.restype = None
    L.mp3dec_b200_version.restype = C.c_char_p
    L.mp3dec_ex_open_buf.argtypes = [C.POINTER(Mp3DecEx), vp, C.c_size_t, C.c_int]
    L.mp3dec_ex_open.argtypes = [C.POINTER(Mp3DecEx), C.c_char_p, C.c_int]
    L.mp3dec_ex_close.argtypes = [C.POINTER(Mp3DecEx)]
    L.mp3dec_ex_close.restype = None
    L.mp3dec_ex_seek.argtypes = [C.POINTER(Mp3DecEx), C.c_uint64]
    L.mp3dec_ex_read.argtypes = [C.POINTER(Mp3DecEx), vp, C.c_size_t]
    L.mp3dec_ex_read.restype = C.c_size_t
    L.mp3dec_load.argtypes = [C.POINTER(Mp3Dec), C.c_char_p, C.POINTER(FileInfo), vp, vp]
    L.mp3dec_detect.argtypes = [C.c_char_p]
    L.free = C.CDLL(None).free
    L.free.argtypes = [vp]
    _lib = L
    return L


def _as_u8(data):
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8)
    return np.frombuffer(bytes(data), dtype=np.uint8) if len(data) else np.zeros(1, np.uint8)


def _opts(device=0, float_output=False, raw_frames=False, allow_mono_stereo=False, host_threads=0, output=OUT_PINNED,
          cuda_stream=None, devices=None, gpu_frame_scan=0, pinned_inputs=None):
    o = BatchOpts(int(device), int(bool(float_output)), int(bool(raw_frames)), int(bool(allow_mono_stereo)),
                  int(host_threads), int(output), C.c_void_p(cuda_stream) if cuda_stream else None)
    o.gpu_frame_scan = int(gpu_frame_scan)
    if pinned_inputs is not None:       # (address, bytes) of a page-locked region holding the stream buffers
        o.pinned_inputs, o.pinned_inputs_bytes = C.c_void_p(int(pinned_inputs[0])), int(pinned_inputs[1])
    if devices:
        o.n_devices = len(devices)
        for i, d in enumerate(devices[:8]):
            o.devices[i] = int(d)
    return o


def shard_streams(sizes, n_shards):
    """Greedy LPT split of streams over GPUs (no collective on the data path)."""
    L = load()
    n = len(sizes)
    arr = (C.c_size_t*n)(*[int(s) for s in sizes])
    out = (C.c_int*n)()
    L.mp3dec_batch_shard(arr, n, int(n_shards), out)
    return list(out)


class BatchPlan:
    """Split form of mp3dec_decode_batch_cuda: parse + H2D at construction, ``run()`` launches the kernels,
    ``fetch()`` copies PCM back.  Keeps references to the input buffers alive."""

    def __init__(self, streams, **kw):
        self.L = load()
        self.float_output = bool(kw.get("float_output", False))
        self._keep = [_as_u8(s) for s in streams]
        n = len(self._keep)
        self.n = n
        self._streams = (BatchStream*max(1, n))()
        for i, (a, s) in enumerate(zip(self._keep, streams)):
            self._streams[i].buf = a.ctypes.data
            self._streams[i].size = len(s)
        self._opts = _opts(**kw)
        err = C.c_int(0)
        self.handle = self.L.mp3dec_batch_plan_create(self._streams, n, C.byref(self._opts), C.byref(err))
        if not self.handle:
            raise RuntimeError("mp3dec_batch_plan_create failed: %d" % err.value)
        st = BatchStats()
        self.L.mp3dec_batch_plan_stats(self.handle, C.byref(st))
        self.stats = {f[0]: getattr(st, f[0]) for f in BatchStats._fields_}

    def enable_taps(self):
        r = self.L.mp3dec_batch_plan_enable_taps(self.handle)
        if r:
            raise RuntimeError("enable_taps failed: %d" % r)

    def run(self):
        r = self.L.mp3dec_batch_plan_run(self.handle)
        if r:
            raise RuntimeError("mp3dec_batch_plan_run failed: %d" % r)

    def run_stage(self, stage):
        r = self.L.mp3dec_batch_plan_run_stage(self.handle, int(stage))
        if r:
            raise RuntimeError("mp3dec_batch_plan_run_stage failed: %d" % r)

    def sync(self):
        r = self.L.mp3dec_batch_plan_sync(self.handle)
        if r:
            raise RuntimeError("mp3dec_batch_plan_sync failed: %d" % r)

    def fetch(self, copy=True):
        """Returns (rets, list of per-stream dict(pcm ndarray, channels, hz, layer, avg_bitrate_kbps))."""
        infos = (FileInfo*max(1, self.n))()
        rets = (C.c_int*max(1, self.n))()
        r = self.L.mp3dec_batch_plan_fetch(self.handle, infos, rets)
        if r:
            raise RuntimeError("mp3dec_batch_plan_fetch failed: %d" % r)
        out = []
        dt = np.float32 if self.float_output else np.int16
        ct = C.c_float if self.float_output else C.c_int16
        for i in range(self.n):
            fi = infos[i]
            if fi.buffer and fi.samples and self._opts.output != OUT_DEVICE:
                arr = np.ctypeslib.as_array(C.cast(fi.buffer, C.POINTER(ct)), shape=(fi.samples,))
                if copy or self._opts.output == OUT_MALLOC:
                    arr = arr.copy()
                if self._opts.output == OUT_MALLOC:
                    self.L.free(fi.buffer)
            else:
                arr = np.zeros(0, dt)
            out.append(dict(pcm=arr, samples=int(fi.samples), channels=fi.channels, hz=fi.hz, layer=fi.layer,
                            avg_bitrate_kbps=fi.avg_bitrate_kbps, device_ptr=fi.buffer if self._opts.output == OUT_DEVICE else None))
        return list(rets)[:self.n], out

    def read_tap(self, which):
        n = int(self.stats["n_granule_channels"])
        shapes = {0: ((n, 576), np.int16), 1: ((n, 40), np.float32), 2: ((n, 576), np.float32), 3: ((n, 576), np.float32),
                  4: ((n, 32), np.uint8), 5: ((n, 40), np.uint8), 6: ((n,), np.uint16),
                  7: ((int(self.stats["n_tiles"]), 16), np.int64)}
        if which == 8:
            # lane order of the entropy kernel: one entry per real Layer III granule-channel (<= n), rest stays ~0
            a = np.full(n, 0xFFFFFFFF, np.uint32)
            r = self.L.mp3dec_batch_plan_read_tap(self.handle, which, a.ctypes.data, a.nbytes)
            if r:
                raise RuntimeError("read_tap(8) failed: %d" % r)
            return a[a != 0xFFFFFFFF]
        shape, dt = shapes[which]
        a = np.zeros(shape, dt)
        r = self.L.mp3dec_batch_plan_read_tap(self.handle, which, a.ctypes.data, a.nbytes)
        if r:
            raise RuntimeError("read_tap(%d) failed: %d" % (which, r))
        return a

    def close(self):
        if getattr(self, "handle", None):
            self.L.mp3dec_batch_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _Batch:
    """ctypes argument pack for the one-call entry point (reusable across calls for benchmarking)."""

    def __init__(self, streams, **kw):
        self.L = load()
        self.float_output = bool(kw.get("float_output", False))
        self.keep = [_as_u8(s) for s in streams]
        self.n = len(self.keep)
        self.streams = (BatchStream*max(1, self.n))()
        for i, (a, s) in enumerate(zip(self.keep, streams)):
            self.streams[i].buf = a.ctypes.data
            self.streams[i].size = len(s)
        self.opts = _opts(**kw)
        self.infos = (FileInfo*max(1, self.n))()
        self.rets = (C.c_int*max(1, self.n))()

    def call(self):
        err = self.L.mp3dec_decode_batch_cuda(self.streams, self.n, self.infos, self.rets, C.byref(self.opts))
        if err:
            raise RuntimeError("mp3dec_decode_batch_cuda failed: %d" % err)
        if not self.n:
            return 0
        # mp3dec_file_info_t.samples of every stream in one go (a Python loop over 8192 ctypes structs costs ms)
        words = np.frombuffer(self.infos, dtype=np.uint8).reshape(self.n, C.sizeof(FileInfo))
        off = FileInfo.samples.offset
        return int(words[:, off:off + 8].copy().view(np.uint64).sum())

    def call_stats(self):
        """where the time of the last call() on this thread went (mp3dec_batch_last_call_stats)"""
        cs = BatchCallStats()
        self.L.mp3dec_batch_last_call_stats(C.byref(cs))
        return {f[0]: getattr(cs, f[0]) for f in BatchCallStats._fields_}

    def free_outputs(self):
        """MP3D_BATCH_OUT_MALLOC: hand the per-stream buffers back (the caller owns them)"""
        if self.opts.output == OUT_MALLOC:
            for i in range(self.n):
                if self.infos[i].buffer:
                    self.L.free(self.infos[i].buffer)
                    self.infos[i].buffer = None

    def results(self, copy=True):
        dt = np.float32 if self.float_output else np.int16
        ct = C.c_float if self.float_output else C.c_int16
        out = []
        for i in range(self.n):
            fi = self.infos[i]
            if fi.buffer and fi.samples and self.opts.output != OUT_DEVICE:
                arr = np.ctypeslib.as_array(C.cast(fi.buffer, C.POINTER(ct)), shape=(fi.samples,))
                if copy or self.opts.output == OUT_MALLOC:
                    arr = arr.copy()
                if self.opts.output == OUT_MALLOC:
                    self.L.free(fi.buffer)
            else:
                arr = np.zeros(0, dt)
            out.append(dict(pcm=arr, samples=int(fi.samples), channels=fi.channels, hz=fi.hz, layer=fi.layer,
                            avg_bitrate_kbps=fi.avg_bitrate_kbps))
        return list(self.rets)[:self.n], out


def mp3dec_decode_batch_cuda(streams, **kw):
    """Decode a list of byte buffers through the C entry point mp3dec_decode_batch_cuda().  Per-stream results
    follow mp3dec_load_buf (minimp3_ex.h:308-525).  Returns (rets, list of dict(pcm, samples, channels, ...))."""
    b = _Batch(streams, **kw)
    b.call()
    return b.results(copy=True)


PROGRESS_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_size_t, C.c_uint64, C.POINTER(FrameInfo))


def mp3dec_load_buf(data, float_output=False, progress=None):
    """mp3dec_load_buf through the C entry point; returns (ret, pcm, info dict).

    ``progress(file_size, offset, frame_info) -> int`` mirrors MP3D_PROGRESS_CB (minimp3_ex.h:52): called once per
    decoded frame, a non-zero return stops the decode and becomes ``ret``."""
    L = load()
    dec = Mp3Dec()
    fi = FileInfo()
    buf = _as_u8(data)
    fn = L.mp3dec_load_buf_f32 if float_output else L.mp3dec_load_buf
    cb = PROGRESS_CB(lambda ud, size, off, info: int(progress(size, off, info.contents))) if progress else None
    ret = fn(C.byref(dec), buf.ctypes.data, len(data), C.byref(fi), cb, None)
    ct = C.c_float if float_output else C.c_int16
    if fi.buffer and fi.samples:
        pcm = np.ctypeslib.as_array(C.cast(fi.buffer, C.POINTER(ct)), shape=(fi.samples,)).copy()
    else:
        pcm = np.zeros(0, np.float32 if float_output else np.int16)
    if fi.buffer:
        L.free(fi.buffer)
    return ret, pcm, dict(channels=fi.channels, hz=fi.hz, layer=fi.layer, avg_bitrate_kbps=fi.avg_bitrate_kbps)


def mp3dec_detect_buf(data):
    L = load()
    buf = _as_u8(data)
    return L.mp3dec_detect_buf(buf.ctypes.data, len(data))


def decode_frames(data, float_output=False):
    """The README's frame loop over mp3dec_init / mp3dec_decode_frame.  Returns (pcm, frames)."""
    L = load()
    dec = Mp3Dec()
    L.mp3dec_init(C.byref(dec))
    info = FrameInfo()
    buf = _as_u8(data)
    dt = np.float32 if float_output else np.int16
    tmp = np.zeros(MINIMP3_MAX_SAMPLES_PER_FRAME, dt)
    fn = L.mp3dec_decode_frame_f32 if float_output else L.mp3dec_decode_frame
    pos, out, frames = 0, [], []
    n = len(data)
    while True:
        samples = fn(C.byref(dec), buf.ctypes.data + pos, min(n - pos, 2**31 - 1), tmp.ctypes.data, C.byref(info))
        frames.append(dict(offset=pos, frame_bytes=info.frame_bytes, frame_offset=info.frame_offset, samples=samples,
                           channels=info.channels, hz=info.hz, layer=info.layer, bitrate_kbps=info.bitrate_kbps))
        if samples:
            out.append(tmp[:samples*info.channels].copy())
        pos += info.frame_bytes
        if not info.frame_bytes:
            break
    return (np.concatenate(out) if out else np.zeros(0, dt)), frames
