"""The iterate / streaming / seek entry points (minimp3_ex.h:104-118) of the drop-in library against the reference:
same stream length, same number of samples delivered after a seek, PCM within +-1 LSB."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_vector, vector_path

pytestmark = pytest.mark.gpu

SEEK_TO_SAMPLE, ALLOW_TRANSITION = 1, 4


@pytest.fixture(scope="module")
def m():
    import minimp3_b200
    minimp3_b200.load()
    return minimp3_b200


def ref_ex(ref, data, flags, seek, portion):
    L = ref.lib
    L.ref_ex_decode.restype = C.c_long
    L.ref_ex_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint64, C.c_size_t, C.c_void_p, C.c_long,
                                C.POINTER(C.c_uint64), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                C.POINTER(C.c_int)]
    buf = np.frombuffer(data, dtype=np.uint8)
    out = np.zeros(4_000_000, np.int16)
    tot, oret, lerr, ch, hz = C.c_uint64(0), C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    n = L.ref_ex_decode(buf.ctypes.data, len(buf), flags, seek, portion, out.ctypes.data, len(out), C.byref(tot),
                        C.byref(oret), C.byref(lerr), C.byref(ch), C.byref(hz))
    return out[:n].copy(), tot.value, oret.value, lerr.value, ch.value, hz.value


def our_ex(m, data, flags, seek, portion):
    L = m.load()
    buf = np.frombuffer(data, dtype=np.uint8)
    dec = m.Mp3DecEx()
    oret = L.mp3dec_ex_open_buf(C.byref(dec), buf.ctypes.data, len(buf), flags)
    if oret:
        return np.zeros(0, np.int16), 0, oret, 0, 0, 0
    tot, ch, hz = dec.samples, dec.info.channels, dec.info.hz
    if seek:
        assert L.mp3dec_ex_seek(C.byref(dec), seek) == 0
    out = np.zeros(4_000_000, np.int16)
    got = 0
    while got + portion <= len(out):
        n = L.mp3dec_ex_read(C.byref(dec), out.ctypes.data + 2*got, portion)
        got += n
        if n != portion:
            break
    lerr = dec.last_error
    L.mp3dec_ex_close(C.byref(dec))
    return out[:got].copy(), tot, oret, lerr, ch, hz


@pytest.mark.parametrize("name,seek", [("l3-compl", 0), ("l3-compl", 100001), ("l3-nonstandard-sin1k0db_lame_vbrtag", 0),
                                       ("l3-nonstandard-sin1k0db_lame_vbrtag", 250000), ("l3-he_mode", 77777),
                                       ("l2-fl10", 5000), ("l3-nonstandard-id3v2", 0), ("M2L3_bitrate_16_all", 123456),
                                       ("l3-si_block", 73000)])
def test_stream_read_and_seek(name, seek, m, ref):
    data = load_vector(name)
    flags = SEEK_TO_SAMPLE | ALLOW_TRANSITION
    want = ref_ex(ref, data, flags, seek, 1152*3 + 7)
    got = our_ex(m, data, flags, seek, 1152*3 + 7)
    assert got[1:] == want[1:], (got[1:], want[1:])            # stream length, open result, last_error, channels, hz
    assert len(got[0]) == len(want[0])
    if len(want[0]):
        assert np.abs(got[0].astype(np.int32) - want[0].astype(np.int32)).max() <= 1


def test_seek_to_byte(m, ref):
    data = load_vector("l3-sin1k0db")
    want = ref_ex(ref, data, 0, 633, 4096)
    got = our_ex(m, data, 0, 633, 4096)
    assert got[1:] == want[1:] and len(got[0]) == len(want[0])
    assert np.abs(got[0].astype(np.int32) - want[0].astype(np.int32)).max() <= 1


def test_iterate_buf_offsets(m, ref):
    data = load_vector("l3-nonstandard-id3v2")
    L = m.load()
    ITER = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_uint64, C.c_void_p)
    seen = []
    cb = ITER(lambda user, frame, size, ff, bufsize, off, info: seen.append((off, size)) or 0)
    buf = np.frombuffer(data, dtype=np.uint8)
    L.mp3dec_iterate_buf.argtypes = [C.c_void_p, C.c_size_t, ITER, C.c_void_p]
    assert L.mp3dec_iterate_buf(buf.ctypes.data, len(buf), cb, None) == 0
    offs = (C.c_uint64*4096)()
    sizes = (C.c_int*4096)()
    ref.lib.ref_iterate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
    n = ref.lib.ref_iterate(buf.ctypes.data, len(buf), offs, sizes, 4096)
    assert seen == [(offs[i], sizes[i]) for i in range(n)] and n > 0


def test_stdio_flavours(m, ref):
    path = vector_path("l3-si_huff").encode()
    L = m.load()
    assert L.mp3dec_detect(path) == 0
    dec, fi = m.Mp3Dec(), m.FileInfo()
    assert L.mp3dec_load(C.byref(dec), path, C.byref(fi), None, None) == 0
    pcm = np.ctypeslib.as_array(C.cast(fi.buffer, C.POINTER(C.c_int16)), shape=(fi.samples,)).copy()
    L.free(fi.buffer)
    want_ret, want, info = ref.load_buf(load_vector("l3-si_huff"))
    assert len(pcm) == len(want) and np.abs(pcm.astype(np.int32) - want.astype(np.int32)).max() <= 1
    ex = m.Mp3DecEx()
    assert L.mp3dec_ex_open(C.byref(ex), path, SEEK_TO_SAMPLE) == 0
    assert ex.samples == len(want)
    L.mp3dec_ex_close(C.byref(ex))
    assert L.mp3dec_ex_open(C.byref(ex), b"/nonexistent/file.mp3", SEEK_TO_SAMPLE) == m.MP3D_E_IOERROR


@pytest.mark.parametrize("name", ["l3-compl", "l3-nonstandard-sin1k0db_lame_vbrtag", "l3-nonstandard-id3v2", "l1-fl1",
                                  "l2-fl10"])
def test_load_buf_progress_callback(name, m, ref):
    """MP3D_PROGRESS_CB (minimp3_ex.h:52, 504-509): one call per decoded frame with a growing offset; the result is
    the one of the callback-free call; a non-zero return cancels and is handed back."""
    data = load_vector(name)
    rret, rpcm, rinfo = ref.load_buf(data)
    calls = []
    ret, pcm, info = m.mp3dec_load_buf(data, progress=lambda size, off, fi: calls.append((size, off, fi.frame_bytes, fi.hz)) or 0)
    assert ret == rret and info == rinfo and len(pcm) == len(rpcm)
    assert np.abs(pcm.astype(np.int32) - rpcm.astype(np.int32)).max() <= 1
    assert calls and all(c[0] == len(data) for c in calls) and calls[-1][1] <= len(data)
    assert all(b[1] > a[1] for a, b in zip(calls, calls[1:])) and all(c[3] == rinfo["hz"] for c in calls)
    ret0, pcm0, _ = m.mp3dec_load_buf(data)                     # the one-shot batch flavour gives the same stream
    assert ret0 == ret and len(pcm0) == len(pcm) and np.abs(pcm0.astype(np.int32) - pcm.astype(np.int32)).max() <= 1
    if len(calls) <= 3:
        return                                                  # one-frame vectors: nothing left to cancel
    seen = []
    ret, part, _ = m.mp3dec_load_buf(data, progress=lambda size, off, fi: seen.append(off) or (5 if len(seen) == 3 else 0))
    assert ret == 5 and len(seen) == 3 and 0 < len(part) < len(pcm)
    assert np.array_equal(part, pcm[:len(part)])


def test_plain_c_caller(tmp_path, manifest, ref):
    """examples/decode_file.c (C99, only the public header): mp3dec_load_buf and the README frame loop."""
    import subprocess
    from test_abi import build_c_example
    exe = build_c_example(str(tmp_path))
    name = "l3-compl"
    out = str(tmp_path / "out.pcm")
    line = subprocess.check_output([exe, vector_path(name), out]).decode()
    kv = dict(t.split("=") for t in line.split())
    mf = manifest[name]
    assert int(kv["samples"]) == mf["load_samples"] and int(kv["frame_loop_samples"]) == mf["samples"]
    assert int(kv["rate"]) == mf["hz"] and int(kv["layer"]) == 3
    pcm = np.fromfile(out, dtype="<i2")
    _, want, _ = ref.load_buf(load_vector(name))
    assert len(pcm) == len(want) and np.abs(pcm.astype(np.int32) - want.astype(np.int32)).max() <= 1


def test_plain_c_batch_caller(tmp_path, manifest):
    """examples/decode_batch.c: per file what mp3dec_load_buf reports (return code, format, sample count)"""
    import subprocess
    from test_abi import build_c_example
    exe = build_c_example(str(tmp_path), "decode_batch")
    names = ["l3-compl", "l3-nonstandard-sin1k0db_lame_vbrtag", "l2-fl10", "l3-compl", "l1-fl1"]
    lines = subprocess.check_output([exe] + [vector_path(n) for n in names]).decode().strip().splitlines()
    assert len(lines) == len(names)
    for n, line in zip(names, lines):
        kv = dict(t.split("=") for t in line.split()[1:])
        mf = manifest[n]
        assert int(kv["ret"]) == mf["load_ret"] and int(kv["samples"]) == mf["load_samples"] and int(kv["rate"]) == mf["hz"]
    assert lines[0].split()[1:] == lines[3].split()[1:]          # the same file twice decodes to the same PCM


@pytest.mark.parametrize("name,window", [("l3-compl", 7), ("l3-he_mode", 20), ("l3-sin1k0db", 33), ("M2L3_bitrate_16_all", 16),
                                         ("l2-fl10", 5), ("l3-test45", 50)])
def test_streaming_across_look_ahead_windows(name, window, m, ref, monkeypatch):
    """mp3dec_ex_read decodes look-ahead windows in one batch each; a window that starts mid-stream re-decodes 16
    frames for the decoder state.  With tiny windows (many boundaries) the stream is still the reference's."""
    monkeypatch.setenv("MP3B_DEBUG_WINDOW_FRAMES", str(window))
    data = load_vector(name)
    for portion in (1000, 1152, 3463):          # the reference's end-of-stream error depends on how reads align with frames
        want, tot_r, oret_r, lerr_r, ch_r, hz_r = ref_ex(ref, data, SEEK_TO_SAMPLE | ALLOW_TRANSITION, 0, portion)
        got, tot, oret, lerr, ch, hz = our_ex(m, data, SEEK_TO_SAMPLE | ALLOW_TRANSITION, 0, portion)
        assert (tot, oret, lerr, ch, hz) == (tot_r, oret_r, lerr_r, ch_r, hz_r) and len(got) == len(want), portion
        assert np.abs(got.astype(np.int32) - want.astype(np.int32)).max() <= 1
    # and after a seek into the middle
    pos = (len(want)//3) & ~1
    want2 = ref_ex(ref, data, SEEK_TO_SAMPLE | ALLOW_TRANSITION, pos, 777)[0]
    got2 = our_ex(m, data, SEEK_TO_SAMPLE | ALLOW_TRANSITION, pos, 777)[0]
    assert len(got2) == len(want2) and np.abs(got2.astype(np.int32) - want2.astype(np.int32)).max() <= 1


def test_streaming_speed_uses_windows(m):
    """whole-stream read through mp3dec_ex_read: far less than one GPU round trip (~50 us) per frame"""
    import time
    data = load_vector("l3-sin1k0db")
    our_ex(m, data, SEEK_TO_SAMPLE, 0, 4096)
    t0 = time.perf_counter()
    got = our_ex(m, data, SEEK_TO_SAMPLE, 0, 4096)[0]
    dt = time.perf_counter() - t0
    assert len(got) == 725760 and dt < 0.010*3, dt
