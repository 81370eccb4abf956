"""Round-2 parity additions (B200): BASELINE.json configs 3, 4, 5 and the dense stress stream at their full sizes,
float output on every Layer III / Layer I-II conformance vector, exactness of damaged streams away from the damage,
a mixed 32-stream window bit for bit, the caller's mp3dec_t copied / restored mid-stream, look-ahead windows at the
lowest bit rates, multi-device sharding inside the batch call."""
import ctypes as C
import zlib

import numpy as np
import pytest

from conftest import L3_CONFORMANCE, L12_CONFORMANCE, load_vector

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def m():
    import minimp3_b200
    minimp3_b200.load()
    return minimp3_b200


def real_rows(desc):
    body = desc[:, :28].astype(np.int64).sum(axis=1)
    return np.where(~((body == 0) & (desc[:, 29] == 1)))[0]


# ---- full-size configs -------------------------------------------------------------------------------------

@pytest.mark.parametrize("config", ["3", "4", "5", "mips", "2b"])
def test_full_size_config_every_stream_equals_its_single_decode(config, m, ref):
    """Every stream of the full-size batch (3072 / 5120 / 8192 / 1024 streams, each with its own host buffer) must be
    byte-identical to the single-stream decode of its vector (the batch is invisible), which in turn is within 1 LSB
    of the unmodified reference; sample totals must add up."""
    from minimp3_b200 import corpus
    streams, names = corpus.build_streams(config)
    distinct = sorted(set(names))
    single = {}
    for n in distinct:
        data = corpus.load_vector_file(n)
        rets, outs = m.mp3dec_decode_batch_cuda([data], allow_mono_stereo=True)
        want_ret, want, _ = ref.load_buf(data)
        assert rets[0] == want_ret == 0, n
        got = outs[0]["pcm"]
        assert len(got) == len(want), n
        d = np.abs(got.astype(np.int32) - want.astype(np.int32))
        assert d.max() <= 1 and (d != 0).mean() < 0.01, n
        single[n] = got
    b = m._Batch(streams, allow_mono_stereo=True, output=m.OUT_PINNED)
    total = b.call()
    rets, outs = b.results(copy=False)
    assert all(r == 0 for r in rets)
    assert total == sum(len(single[n]) for n in names)
    bad = [i for i, (n, o) in enumerate(zip(names, outs)) if not np.array_equal(o["pcm"], single[n])]
    assert not bad, (len(bad), bad[:5])
    m.load().mp3dec_batch_release_outputs(0)


def test_multi_device_call_matches_single_device(m):
    """opts.devices: the shards of one call (here: all on device 0, listed several times) give every stream the bytes
    of the plain call, in the caller's order"""
    names = ["l3-si_block", "l3-compl", "l3-he_mode", "M2L3_noise", "l3-si_huff", "l3-he_free", "l1-fl1", "l2-fl10",
             "l3-sin1k0db", "l3-nonstandard-small"]
    datas = [load_vector(n) for n in names]*7 + [b""]
    rets0, whole = m.mp3dec_decode_batch_cuda(datas, allow_mono_stereo=True, output=m.OUT_MALLOC)
    for devs in ([0, 0], [0, 0, 0, 0, 0, 0, 0, 0]):
        rets, outs = m.mp3dec_decode_batch_cuda(datas, allow_mono_stereo=True, output=m.OUT_MALLOC, devices=devs)
        assert rets == rets0
        for a, b in zip(whole, outs):
            assert np.array_equal(a["pcm"], b["pcm"]) and a["hz"] == b["hz"] and a["channels"] == b["channels"]


# ---- tightened parity ---------------------------------------------------------------------------------------

@pytest.mark.parametrize("name", L3_CONFORMANCE + L12_CONFORMANCE)
def test_float_output_within_1e5_all_vectors(name, m, ref_f32):
    data = load_vector(name)
    want, _ = ref_f32.decode_frames(data)
    rets, outs = m.mp3dec_decode_batch_cuda([data], raw_frames=True, float_output=True)
    got = outs[0]["pcm"]
    assert got.dtype == np.float32 and len(got) == len(want)
    assert np.abs(got - want).max() <= 1e-5


def test_mixed_window_is_bit_exact_on_the_integers(m, ref):
    """32 different streams (sample rates, block types, mono / stereo, MPEG-1/2/2.5) side by side in one lane window
    of the entropy kernel, several windows deep: quantised integers and scalefactors bit for bit, PCM within 1 LSB"""
    # (l3-nonstandard-big-iscf is left out: its scalefactors push the spectrum to ~1e9, where the PCM is a difference
    # of huge floats and any other operation order moves it by tens of LSB; its integers are covered elsewhere)
    names = (L3_CONFORMANCE + ["l3-nonstandard-compl-sideinfo-bigvalues", "l3-nonstandard-he_44_48khz",
                               "l3-nonstandard-compl-sideinfo-blocktype"] + L3_CONFORMANCE[:11])[:32]
    assert len(names) == 32
    datas = [load_vector(n) for n in names]*9          # > 8192 granule-channels: the cross-stream lane order engages
    taps = {n: ref.tap_stream(load_vector(n)) for n in set(names)}
    p = m.BatchPlan(datas, raw_frames=True)
    p.enable_taps()
    p.run()
    rets, outs = p.fetch()
    desc, q, scf = p.read_tap(4), p.read_tap(0), p.read_tap(1)
    p.close()
    rows = real_rows(desc)
    at = 0
    for n, o in zip(names*9, outs):
        t = taps[n]
        k = len(t["rec"])
        r = rows[at:at + k]
        at += k
        assert np.array_equal(q[r], t["q"]), n
        for i, rec in enumerate(t["rec"]):
            nb = rec.n_long_sfb + rec.n_short_sfb
            assert np.array_equal(scf[r[i], :nb].view(np.uint32), t["scf"][i, :nb].view(np.uint32)), n
        assert len(o["pcm"]) == len(t["pcm"])
        assert np.abs(o["pcm"].astype(np.int32) - t["pcm"].astype(np.int32)).max() <= 1, n
    assert at == len(rows)


def test_damaged_streams_are_exact_away_from_the_damage(m, ref):
    """byte damage inside the main data of chosen frames (headers and side info untouched, so the frame structure and
    every reservoir pointer stay valid): PCM must be within 1 LSB of the reference everywhere except in the frames
    whose bits were hit -- the damaged frame, the frames whose main data reach back into it, and one frame of
    overlap / synthesis memory behind them"""
    import harness
    for name in ["l3-sin1k0db", "l3-compl", "M2L3_bitrate_22_all", "l3-he_44khz"]:
        base = load_vector(name)
        _, frames = ref.decode_frames(base)
        rng = np.random.default_rng(zlib.crc32(name.encode()))
        good = [f for f in frames if f.samples and f.frame_bytes > 80]
        d = bytearray(base)
        hit = sorted(rng.choice(len(good) - 8, 5, replace=False) + 4)
        for k in hit:
            f = good[k]
            lo = f.offset + f.frame_offset + 4 + 2 + 32 + 4          # behind header, CRC and any side info
            hi = f.offset + f.frame_bytes
            for _ in range(6):
                d[int(rng.integers(lo, hi))] ^= int(rng.integers(1, 256))
        data = bytes(d)
        want, wframes = ref.decode_frames(data)
        rets, outs = m.mp3dec_decode_batch_cuda([data], raw_frames=True)
        got = outs[0]["pcm"]
        assert rets[0] == 0 and len(got) == len(want)
        assert [(f.samples, f.frame_bytes) for f in wframes] == [(f.samples, f.frame_bytes) for f in frames]
        # sample ranges that may legitimately differ: from each damaged frame to 4 frames behind it
        ok = np.ones(len(want), bool)
        pos, starts = 0, {}
        for f in wframes:
            starts[f.offset] = pos
            pos += f.samples*f.channels
        order = [f for f in wframes if f.samples]
        for k in hit:
            i = order.index(next(f for f in order if f.offset == good[k].offset))
            a = starts[order[max(0, i - 1)].offset]
            b = starts[order[i + 5].offset] if i + 5 < len(order) else len(want)
            ok[a:b] = False
        diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
        assert diff[ok].max() <= 1, name
        assert ok.mean() > 0.7


# ---- the caller's mp3dec_t is the authority (advisor finding, round 1) ----------------------------------------

def _loop(m, L, dec, buf, pos, n_calls, out):
    info = m.FrameInfo()
    tmp = np.zeros(m.MINIMP3_MAX_SAMPLES_PER_FRAME, np.int16)
    for _ in range(n_calls):
        s = L.mp3dec_decode_frame(C.byref(dec), buf.ctypes.data + pos, len(buf) - pos, tmp.ctypes.data, C.byref(info))
        if s:
            out.append(tmp[:s*info.channels].copy())
        pos += info.frame_bytes
        if not info.frame_bytes:
            return pos, True
    return pos, False


@pytest.mark.parametrize("name", ["l3-sin1k0db", "l3-he_mode", "M2L3_bitrate_22_all"])
def test_decoder_struct_copied_and_restored_mid_stream(name, m, ref):
    L = m.load()
    data = load_vector(name)
    want, _ = ref.decode_frames(data)
    buf = np.frombuffer(data, dtype=np.uint8).copy()
    # (a) the struct moves to another address after 40 frames (a C++ wrapper holding it by value)
    a = m.Mp3Dec()
    L.mp3dec_init(C.byref(a))
    out = []
    pos, done = _loop(m, L, a, buf, 0, 40, out)
    b = m.Mp3Dec()
    C.memmove(C.byref(b), C.byref(a), C.sizeof(a))
    C.memset(C.byref(a), 0x5a, C.sizeof(a))          # the old home is gone
    pos, done = _loop(m, L, b, buf, pos, 10**9, out)
    pcm = np.concatenate(out)
    assert done and len(pcm) == len(want) and np.abs(pcm.astype(np.int32) - want.astype(np.int32)).max() <= 1
    # (b) a snapshot taken after 25 frames is restored after 70 and decoding resumes from the snapshot's position
    d = m.Mp3Dec()
    L.mp3dec_init(C.byref(d))
    first = []
    pos25, _ = _loop(m, L, d, buf, 0, 25, first)
    snap = m.Mp3Dec()
    C.memmove(C.byref(snap), C.byref(d), C.sizeof(d))
    n25 = sum(len(x) for x in first)
    _loop(m, L, d, buf, pos25, 45, first)
    C.memmove(C.byref(d), C.byref(snap), C.sizeof(d))
    again = []
    _, done = _loop(m, L, d, buf, pos25, 10**9, again)
    pcm2 = np.concatenate(again)
    assert done and len(pcm2) == len(want) - n25
    assert np.abs(pcm2.astype(np.int32) - want[n25:].astype(np.int32)).max() <= 1


@pytest.mark.parametrize("name", ["M2L3_bitrate_24_all", "M2L3_bitrate_16_all", "M2L3_bitrate_22_all"])
@pytest.mark.parametrize("window", [5, 9, 33])
def test_window_boundaries_at_the_lowest_bit_rates(name, window, m, ref, monkeypatch):
    """8 kbit/s MPEG-2 frames carry ~11 payload bytes while main_data_begin reaches 255 bytes back: a window that
    starts mid-stream has to re-decode by reservoir BYTES, not by a fixed frame count (advisor finding, round 1)"""
    monkeypatch.setenv("MP3B_DEBUG_WINDOW_FRAMES", str(window))
    data = load_vector(name)
    want, _ = ref.decode_frames(data)
    pcm, frames = m.decode_frames(data)
    assert len(pcm) == len(want) and np.abs(pcm.astype(np.int32) - want.astype(np.int32)).max() <= 1
    L = m.load()
    buf = np.frombuffer(data, dtype=np.uint8)
    dec = m.Mp3DecEx()
    assert L.mp3dec_ex_open_buf(C.byref(dec), buf.ctypes.data, len(buf), 1) == 0
    out = np.zeros(len(want) + 4096, np.int16)
    got = 0
    while True:
        n = L.mp3dec_ex_read(C.byref(dec), out.ctypes.data + 2*got, 7001)
        got += n
        if n != 7001:
            break
    L.mp3dec_ex_close(C.byref(dec))
    want_lb = ref.load_buf(data)[1]
    assert got == len(want_lb) and np.abs(out[:got].astype(np.int32) - want_lb.astype(np.int32)).max() <= 1


def test_pinned_outputs_survive_internal_batch_calls(m, ref):
    """MP3D_BATCH_OUT_PINNED pointers stay valid until the caller's next pinned / device call: the library's own
    batch calls in between (mp3dec_load_buf, the frame loop's look-ahead windows) must not recycle the slab"""
    datas = [load_vector("l3-sin1k0db"), load_vector("l3-compl")]*8
    b = m._Batch(datas, raw_frames=True, output=m.OUT_PINNED)
    b.call()
    _, views = b.results(copy=False)
    keep = [v["pcm"].copy() for v in views]
    for _ in range(20):
        m.mp3dec_load_buf(load_vector("l3-he_44khz"))
        m.decode_frames(load_vector("l3-he_mode"))
        m.mp3dec_decode_batch_cuda([load_vector("l3-test45")]*4, output=m.OUT_MALLOC)
    for k, v in zip(keep, views):
        assert np.array_equal(k, v["pcm"])
    m.load().mp3dec_batch_release_outputs(0)


# ---- frame sync + reservoir accounting on the GPU (k_scan.cu) ------------------------------------

SCAN_REGULAR = ["l3-compl", "l3-he_32khz", "l3-he_44khz", "l3-he_48khz", "l3-hecommon", "l3-si", "l3-si_block",
                "l3-si_huff", "l3-sin1k0db", "M2L3_bitrate_16_all", "M2L3_bitrate_22_all", "M2L3_bitrate_24_all",
                "M2L3_compl24", "M2L3_noise", "l3-nonstandard-sin1k0db_lame_vbrtag"]


def _call(m, datas, **kw):
    b = m._Batch(datas, output=m.OUT_MALLOC, **kw)
    b.call()
    st = b.call_stats()
    rets, outs = b.results()
    return rets, outs, st


@pytest.mark.parametrize("raw", [False, True])
def test_gpu_frame_scan_is_invisible(raw, m):
    """one batch through the GPU frame scan (forced) and through the host walker: same return codes, stream
    parameters and PCM bytes for every stream -- regular Layer III streams (scanned), Layer I/II, free format,
    mono/stereo switches, tags, damage, junk and empty buffers (handed back to the host walker or empty)"""
    rng = np.random.default_rng(99)
    names = SCAN_REGULAR + ["l3-he_free", "l3-he_mode", "l3-test45", "l1-fl1", "l2-fl10", "l3-nonstandard-id3v2",
                            "l3-nonstandard-id3v1-apetag", "l3-nonstandard-apetag", "l3-nonstandard-vbrtag-only",
                            "l3-nonstandard-small", "l3-nonstandard-compl-sideinfo-size", "l3-nonstandard-he_44_48khz"]
    datas = [load_vector(n) for n in names]
    base = load_vector("l3-sin1k0db")
    for t in range(12):
        b = bytearray(base)
        if t % 3 == 0:
            for _ in range(4):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        elif t % 3 == 1:
            b = b[int(rng.integers(1, 3000)):int(rng.integers(60000, len(b)))]
        else:
            p, n = int(rng.integers(0, len(b) - 500)), int(rng.integers(1, 500))
            del b[p:p + n]
        datas.append(bytes(b))
    datas += [b"", b"\xff"*777, base[:100], base + base]
    r0, o0, s0 = _call(m, datas, raw_frames=raw, allow_mono_stereo=True, gpu_frame_scan=-1)
    r1, o1, s1 = _call(m, datas, raw_frames=raw, allow_mono_stereo=True, gpu_frame_scan=1)
    assert s0["scanned_streams"] == 0 and s1["scanned_streams"] >= len(SCAN_REGULAR)
    assert r0 == r1
    for i, (a, b) in enumerate(zip(o0, o1)):
        assert (a["samples"], a["channels"], a["hz"], a["layer"], a["avg_bitrate_kbps"]) == \
               (b["samples"], b["channels"], b["hz"], b["layer"], b["avg_bitrate_kbps"]), i
        assert np.array_equal(a["pcm"], b["pcm"]), i


@pytest.mark.parametrize("name", SCAN_REGULAR)
def test_gpu_frame_scan_descriptors_match_host_walk(name, m):
    """the records the scan kernels emit lead to byte-identical granule descriptors (bit offsets into the main data
    included), the same lane-order set and the same PCM as the host walk of the stream"""
    data = load_vector(name)
    res = []
    for scan in (-1, 1):
        p = m.BatchPlan([data, data], raw_frames=True, gpu_frame_scan=scan)
        p.run()
        rets, outs = p.fetch()
        res.append((np.ascontiguousarray(p.read_tap(4)), np.sort(p.read_tap(8)), outs, dict(p.stats)))
        p.close()
    (d0, o0, out0, st0), (d1, o1, out1, st1) = res
    assert st0["n_frames"] == st1["n_frames"] and st0["n_tiles"] == st1["n_tiles"] and st0["samples"] == st1["samples"]
    real = np.where(d0.view(np.uint8).reshape(len(d0), -1).any(axis=1))[0]
    a = d0.view(np.uint8).reshape(len(d0), -1)[real].copy()
    b = d1.view(np.uint8).reshape(len(d1), -1)[real].copy()
    # bit offsets are relative to each route's own main-data layout: compare them relative to the stream's first
    off0 = a[:, :8].copy().view(np.uint64).ravel()
    off1 = b[:, :8].copy().view(np.uint64).ravel()
    half = len(real)//2
    assert np.array_equal(off0[:half] - off0[0], off1[:half] - off1[0])
    assert np.array_equal(a[:, 8:], b[:, 8:])
    assert np.array_equal(o0, o1)
    for x, y in zip(out0, out1):
        assert np.array_equal(x["pcm"], y["pcm"])


def test_pinned_input_slab_uploads_directly_same_results(m):
    """opts.pinned_inputs: streams that sit in one page-locked region are uploaded from where they are (no staging
    copy) by the scan route; same bytes out as the ordinary call"""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from e2e_probe import pinned_slab
    names = SCAN_REGULAR + ["l3-he_free", "l1-fl1", "l3-he_mode", "l3-nonstandard-id3v2"]
    datas = [load_vector(n) for n in names]*5
    views, slab, keep = pinned_slab(datas)
    r0, o0, s0 = _call(m, datas, allow_mono_stereo=True, gpu_frame_scan=-1)
    r1, o1, s1 = _call(m, views, allow_mono_stereo=True, gpu_frame_scan=1, pinned_inputs=slab)
    assert r0 == r1 and s1["scanned_streams"] >= 5*len(SCAN_REGULAR)
    for a, b in zip(o0, o1):
        assert a["samples"] == b["samples"] and np.array_equal(a["pcm"], b["pcm"])


def test_gpu_frame_scan_on_many_damaged_streams(m):
    """200 damaged variants of four streams (byte damage anywhere incl. headers and side info, cuts at both ends,
    deleted and duplicated ranges, junk appended) through the forced GPU scan and through the host walker: identical
    return codes, sample counts and PCM bytes -- whatever the scan cannot vouch for must end up with the walker"""
    rng = np.random.default_rng(20240)
    bases = [load_vector(n) for n in ("l3-sin1k0db", "l3-compl", "M2L3_bitrate_22_all", "l3-he_44khz")]
    datas = []
    for t in range(200):
        b = bytearray(bases[t % 4])
        kind = t % 5
        if kind == 0:
            for _ in range(int(rng.integers(1, 8))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        elif kind == 1:
            b = b[int(rng.integers(0, 2000)):int(rng.integers(len(b)//2, len(b)))]
        elif kind == 2:
            p, n = int(rng.integers(0, len(b) - 600)), int(rng.integers(1, 600))
            if rng.random() < 0.5:
                del b[p:p + n]
            else:
                b[p:p] = b[p:p + n]
        elif kind == 3:
            b += rng.integers(0, 256, int(rng.integers(1, 3000)), dtype=np.uint8).tobytes()
        else:
            p = int(rng.integers(0, len(b) - 4))
            b[p:p + 4] = bytes([0xFF, 0xFB, int(rng.integers(0, 256)), int(rng.integers(0, 256))])
        datas.append(bytes(b))
    for raw in (False, True):
        r0, o0, s0 = _call(m, datas, raw_frames=raw, allow_mono_stereo=True, gpu_frame_scan=-1)
        r1, o1, s1 = _call(m, datas, raw_frames=raw, allow_mono_stereo=True, gpu_frame_scan=1)
        assert r0 == r1 and s1["scanned_streams"] > 20
        for i, (a, b) in enumerate(zip(o0, o1)):
            assert a["samples"] == b["samples"] and a["hz"] == b["hz"] and a["channels"] == b["channels"], (raw, i)
            assert np.array_equal(a["pcm"], b["pcm"]), (raw, i)


def test_damaged_layer12_streams_decode_like_the_reference(m, ref):
    """Layer I / II streams with random byte damage, cuts and duplicated stretches in one batch: return codes and sample
    counts of the reference, clean copies within 1 LSB, and -- a Layer I/II frame stands alone (no bit reservoir) -- most
    samples of the damaged copies too.  Exercises k_l12_scale_info / k_l12_dequant on garbage allocations."""
    rng = np.random.default_rng(1212)
    datas, clean = [], []
    for name in ["l1-fl1", "l1-fl5", "l2-fl10", "l2-fl14", "l2-test32", "ILL2_dual"]:
        base = load_vector(name)
        clean.append(len(datas))
        datas.append(base)
        for t in range(10):
            b = bytearray(base)
            if t % 2 == 0:
                for _ in range(int(rng.integers(1, 12))):
                    p, n = int(rng.integers(0, len(b) - 64)), int(rng.integers(1, 48))
                    b[p:p + n] = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
            else:
                for _ in range(int(rng.integers(1, 4))):
                    p, n = int(rng.integers(0, len(b) - 600)), int(rng.integers(1, 600))
                    if rng.random() < 0.5:
                        del b[p:p + n]
                    else:
                        b[p:p] = b[p:p + n]
            datas.append(bytes(b))
    rets, outs = m.mp3dec_decode_batch_cuda(datas, allow_mono_stereo=True)
    close = 0
    for i, (d, r, o) in enumerate(zip(datas, rets, outs)):
        ret, want, info = ref.load_buf(d)
        assert r == ret and len(o["pcm"]) == len(want), i
        if not len(want):
            close += 1
            continue
        diff = np.abs(o["pcm"].astype(np.int32) - want.astype(np.int32))
        if i in clean:
            assert diff.max() <= 1, i
        close += int((diff <= 1).mean() > 0.9)
    assert close >= len(datas)*0.8
