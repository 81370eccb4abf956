"""Parity of the CUDA path (through the C ABI) with the unmodified reference, on a B200.

Bar (BASELINE.json north_star): Huffman-decoded integer spectra and scalefactors bit-exact; PCM within
+-1 LSB int16 of minimp3's own output; FLOAT_OUTPUT within 1e-5 abs; PSNR > 96 dB on every bundled
Layer III conformance vector.
"""
import ctypes as C
import hashlib

import numpy as np
import pytest

import conftest
from conftest import L3_CONFORMANCE, L3_SMOKE, L12_CONFORMANCE, L12_REGRESSION, load_iso_pcm, load_vector, psnr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def m():
    import minimp3_b200
    minimp3_b200.load()
    return minimp3_b200


def real_rows(desc):
    body = desc[:, :28].astype(np.int64).sum(axis=1)
    return np.where(~((body == 0) & (desc[:, 29] == 1)))[0]


@pytest.mark.parametrize("name", L3_CONFORMANCE)
def test_stage_taps_bit_exact_and_pcm_within_one_lsb(name, m, ref):
    data = load_vector(name)
    t = ref.tap_stream(data)
    p = m.BatchPlan([data], raw_frames=True)
    p.enable_taps()
    p.run()
    rets, outs = p.fetch()
    desc, q, scf, xr, sub, nzl = (p.read_tap(k) for k in (4, 0, 1, 2, 3, 6))
    p.close()
    rows = real_rows(desc)
    assert rets[0] == 0 and len(rows) == len(t["rec"])
    assert np.array_equal(q[rows], t["q"])                                   # integer spectra
    for i, r in enumerate(t["rec"]):
        n = r.n_long_sfb + r.n_short_sfb
        assert np.array_equal(scf[rows[i], :n].view(np.uint32), t["scf"][i, :n].view(np.uint32))   # scalefactors
    xrz = xr[rows].copy()
    for i, g in enumerate(rows):
        xrz[i, nzl[g]:] = 0
    # requantised lines, bitwise.  Plain MS granules keep L/R here (the transform kernel forms L+R / L-R while
    # loading); intensity-stereo granules were already processed in place by k_l3_intensity_stereo.
    ms_plain = (desc[rows, 29] & 0x10) != 0
    want = np.where(ms_plain[:, None], t["xr"], t["xr_stereo"])
    assert np.array_equal(xrz.view(np.uint32), want.view(np.uint32))
    scale = max(1.0, float(np.abs(t["sub"]).max()))
    assert np.abs(sub[rows] - t["sub"]).max() <= 2e-6*scale                  # hybrid filterbank output
    pcm = outs[0]["pcm"]
    assert len(pcm) == len(t["pcm"])
    d = np.abs(pcm.astype(np.int32) - t["pcm"].astype(np.int32))
    assert d.max() <= 1 and (d != 0).mean() < 0.01


@pytest.mark.parametrize("name", L3_SMOKE)
def test_cuda_against_oracle_port(name, m):
    """same inputs through the plain-C restatement (oracle/mp3_oracle.c): integers bit-exact, filterbank tight"""
    from oracle import port
    data = load_vector(name)
    o = port.decode_stream(data, taps=True)
    p = m.BatchPlan([data], raw_frames=True)
    p.enable_taps()
    p.run()
    rets, outs = p.fetch()
    desc, q, sub = p.read_tap(4), p.read_tap(0), p.read_tap(3)
    p.close()
    rows = real_rows(desc)
    assert len(rows) == o["n_gc"]
    assert np.array_equal(q[rows], o["q"])
    scale = max(1.0, float(np.abs(o["sub"]).max()))
    assert np.abs(sub[rows] - o["sub"]).max() <= 2e-6*scale
    d = np.abs(outs[0]["pcm"].astype(np.int32) - o["pcm"].astype(np.int32))
    assert len(d) == len(o["pcm"]) and d.max() <= 1


@pytest.mark.parametrize("name", L3_CONFORMANCE)
def test_conformance_psnr(name, m, manifest):
    """minimp3_test.c:391-428 on the CUDA output: sample-count rule and PSNR >= 96 dB against the ISO PCM"""
    data = load_vector(name)
    rets, outs = m.mp3dec_decode_batch_cuda([data], allow_mono_stereo=True)
    pcm = outs[0]["pcm"]
    iso = load_iso_pcm(name)
    assert rets[0] == 0
    assert len(pcm) in (len(iso), len(iso) + 1152, len(iso) + 2304)
    md, ps = psnr(pcm, iso)
    assert ps >= 96.0 and md <= 2
    assert ps >= manifest[name]["ref_psnr"] - 1.0        # not worse than the reference itself scores here


@pytest.mark.parametrize("name", ["l3-nonstandard-compl-sideinfo-bigvalues", "l3-nonstandard-compl-sideinfo-blocktype",
                                  "l3-nonstandard-compl-sideinfo-size", "l3-nonstandard-he_44_48khz",
                                  "l3-nonstandard-sin1k0db_lame_vbrtag", "l3-nonstandard-id3v1", "l3-nonstandard-id3v2",
                                  "l3-nonstandard-apetag", "l3-nonstandard-vbrtag-only", "l3-nonstandard-small"])
def test_load_buf_semantics_on_regression_vectors(name, m, ref):
    """mp3dec_load_buf through the C entry point: return code, sample count, PCM (+-1 LSB) as the reference"""
    data = load_vector(name)
    want_ret, want_pcm, want_info = ref.load_buf(data)
    L = m.load()
    # the reference test binary allows mono/stereo transitions; none of these vectors has one
    ret, pcm, info = m.mp3dec_load_buf(data)
    assert ret == want_ret
    assert len(pcm) == len(want_pcm)
    if len(pcm):
        assert np.abs(pcm.astype(np.int32) - want_pcm.astype(np.int32)).max() <= 1
    for k in ("hz", "layer", "avg_bitrate_kbps", "channels"):
        assert info[k] == want_info[k], k


@pytest.mark.parametrize("name", ["l3-sin1k0db", "l3-he_mode", "M2L3_bitrate_22_all", "l3-si_block"])
def test_float_output_within_1e5(name, m, ref_f32):
    data = load_vector(name)
    want, _ = ref_f32.decode_frames(data)
    rets, outs = m.mp3dec_decode_batch_cuda([data], raw_frames=True, float_output=True)
    got = outs[0]["pcm"]
    assert got.dtype == np.float32 and len(got) == len(want)
    assert np.abs(got - want).max() <= 1e-5


@pytest.mark.parametrize("name", ["l3-compl", "l3-he_mode", "l3-si_block", "M2L3_compl24", "l3-he_free", "l1-fl4", "l2-fl13",
                                  "l2-fl10", "ILL2_layer3", "l2-nonstandard-fl1_fl2_ff"])
def test_frame_api_matches_reference(name, m, ref):
    """mp3dec_init + mp3dec_decode_frame loop (README usage): per-call return value, frame_bytes, info and PCM"""
    data = load_vector(name)
    want_pcm, want_frames = ref.decode_frames(data)
    pcm, frames = m.decode_frames(data)
    assert len(frames) == len(want_frames)
    for a, b in zip(frames, want_frames):
        assert (a["samples"], a["frame_bytes"]) == (b.samples, b.frame_bytes)
        if b.frame_bytes and b.samples:
            assert (a["frame_offset"], a["channels"], a["hz"], a["layer"], a["bitrate_kbps"]) == \
                   (b.frame_offset, b.channels, b.hz, b.layer, b.bitrate_kbps)
    assert len(pcm) == len(want_pcm)
    assert np.abs(pcm.astype(np.int32) - want_pcm.astype(np.int32)).max() <= 1


def test_frame_api_parse_only(m):
    L = m.load()
    data = np.frombuffer(load_vector("l3-compl"), dtype=np.uint8)
    dec, info = m.Mp3Dec(), m.FrameInfo()
    L.mp3dec_init(C.byref(dec))
    n = L.mp3dec_decode_frame(C.byref(dec), data.ctypes.data, len(data), None, C.byref(info))
    assert n == 1152 and info.hz == 48000 and info.channels == 1 and info.layer == 3 and info.frame_bytes > 0


def test_ragged_batch_matches_single_stream_decodes(m, ref):
    """one call, streams of very different length / format, including empty and garbage buffers"""
    names = ["l3-si_huff", "l3-sin1k0db", "l3-nonstandard-small", "M2L3_bitrate_16_all", "l3-he_mode", "l3-he_free",
             "l3-nonstandard-id3v2-only", "l3-test46"]
    datas = [load_vector(n) for n in names] + [b"", b"\xff"*777]
    rets, outs = m.mp3dec_decode_batch_cuda(datas, raw_frames=True)
    for d, r, o in zip(datas, rets, outs):
        want, _ = ref.decode_frames(d)
        assert r == 0 and len(o["pcm"]) == len(want)
        if len(want):
            assert np.abs(o["pcm"].astype(np.int32) - want.astype(np.int32)).max() <= 1


def test_batch_is_deterministic_and_order_independent(m):
    names = ["l3-si_block", "l3-compl", "l3-he_mode", "M2L3_noise"]
    datas = [load_vector(n) for n in names]
    _, a = m.mp3dec_decode_batch_cuda(datas, raw_frames=True)
    _, b = m.mp3dec_decode_batch_cuda(datas[::-1], raw_frames=True)
    for x, y in zip(a, b[::-1]):
        assert np.array_equal(x["pcm"], y["pcm"])


def test_full_size_batch_properties(m, ref):
    """BASELINE.json configs[1] at full size: 1024 x l3-sin1k0db.  Every stream must produce the same PCM as a
    single-stream decode (checksum of checksums), which in turn is within 1 LSB of the reference."""
    data = load_vector("l3-sin1k0db")
    want, _ = ref.decode_frames(data)
    _, one = m.mp3dec_decode_batch_cuda([data])
    assert np.abs(one[0]["pcm"].astype(np.int32) - want.astype(np.int32)).max() <= 1
    h1 = hashlib.sha256(one[0]["pcm"].tobytes()).hexdigest()
    rets, outs = m.mp3dec_decode_batch_cuda([data]*1024)
    assert all(r == 0 for r in rets)
    assert sum(o["samples"] for o in outs) == 1024*725760
    assert {hashlib.sha256(o["pcm"].tobytes()).hexdigest() for o in outs} == {h1}


def test_mixed_block_and_lsf_batches(m, ref):
    """BASELINE.json configs[2] and [3], 64 streams each, cycled"""
    for names in (["l3-si_block", "l3-he_mode", "l3-si_huff"],
                  ["M2L3_bitrate_16_all", "M2L3_bitrate_22_all", "M2L3_bitrate_24_all", "M2L3_compl24", "l3-he_free"]):
        datas = [load_vector(names[i % len(names)]) for i in range(64)]
        wants = {n: ref.decode_frames(load_vector(n))[0] for n in names}
        rets, outs = m.mp3dec_decode_batch_cuda(datas, raw_frames=True)
        for i, o in enumerate(outs):
            w = wants[names[i % len(names)]]
            assert rets[i] == 0 and len(o["pcm"]) == len(w)
            assert np.abs(o["pcm"].astype(np.int32) - w.astype(np.int32)).max() <= 1


def test_sharding_is_invisible_in_the_output(m):
    """multi-GPU path on one device: decode the shards of n_gpus in {1,2,4,8} separately, same bytes out"""
    names = ["l3-si_block", "l3-compl", "l3-he_mode", "M2L3_noise", "l3-si_huff", "l3-he_free", "l3-si", "l3-hecommon",
             "M2L3_compl24"]
    datas = [load_vector(n) for n in names]
    _, whole = m.mp3dec_decode_batch_cuda(datas, raw_frames=True)
    for n_shards in (2, 4, 8):
        shard_of = m.shard_streams([len(d) for d in datas], n_shards)
        assert set(shard_of) <= set(range(n_shards))
        for s in range(n_shards):
            ids = [i for i, v in enumerate(shard_of) if v == s]
            if not ids:
                continue
            _, part = m.mp3dec_decode_batch_cuda([datas[i] for i in ids], raw_frames=True)
            for i, o in zip(ids, part):
                assert np.array_equal(o["pcm"], whole[i]["pcm"])


@pytest.mark.parametrize("name", L12_CONFORMANCE + L12_REGRESSION)
def test_layer12_pcm_within_one_lsb(name, m, ref):
    """Layer I/II: dequantisation + 12-slot synthesis on the GPU against the reference"""
    data = load_vector(name)
    want, _ = ref.decode_frames(data)
    rets, outs = m.mp3dec_decode_batch_cuda([data], raw_frames=True)
    got = outs[0]["pcm"]
    assert rets[0] == 0 and len(got) == len(want)
    if len(want):
        d = np.abs(got.astype(np.int32) - want.astype(np.int32))
        assert d.max() <= 1 and (d != 0).mean() < 0.01


@pytest.mark.parametrize("name", L12_CONFORMANCE)
def test_layer12_conformance_psnr(name, m, manifest):
    data = load_vector(name)
    rets, outs = m.mp3dec_decode_batch_cuda([data], allow_mono_stereo=True)
    iso = load_iso_pcm(name)
    md, ps = psnr(outs[0]["pcm"], iso)
    assert rets[0] == manifest[name]["load_ret"] and len(outs[0]["pcm"]) == manifest[name]["load_samples"]
    assert ps >= 96.0 and ps >= manifest[name]["ref_psnr"] - 1.0


def test_corrupted_streams_decode_without_faults(m, ref):
    """byte damage + truncation: same sample counts as the reference, no CUDA error, and -- away from the damage --
    mostly the same PCM (the reference itself reads stale stack bytes on some malformed granules, so exact
    equality is only required of the undamaged streams in the same batch)"""
    import zlib
    names = ["l3-compl", "l3-he_mode", "l3-si_block", "M2L3_bitrate_22_all", "l2-fl10", "l3-sin1k0db"]
    datas, clean = [], []
    for n in names:
        base = load_vector(n)
        rng = np.random.default_rng(zlib.crc32(n.encode()))
        for trial in range(4):
            d = bytearray(base)
            for _ in range(10):
                d[int(rng.integers(0, len(d)))] = int(rng.integers(0, 256))
            if trial % 2:
                d = d[:int(rng.integers(len(d)//2, len(d)))]
            datas.append(bytes(d))
            clean.append(False)
        datas.append(base)
        clean.append(True)
    rets, outs = m.mp3dec_decode_batch_cuda(datas, raw_frames=True)
    close = 0
    for d, c, r, o in zip(datas, clean, rets, outs):
        want, _ = ref.decode_frames(d)
        assert r == 0 and len(o["pcm"]) == len(want)
        if len(want):
            diff = np.abs(o["pcm"].astype(np.int32) - want.astype(np.int32))
            if c:
                assert diff.max() <= 1
            close += int((diff <= 1).mean() > 0.9)
    assert close >= len(datas)*0.8


def test_layer12_float_output(m, ref_f32):
    data = load_vector("l2-fl12")
    want, _ = ref_f32.decode_frames(data)
    rets, outs = m.mp3dec_decode_batch_cuda([data], raw_frames=True, float_output=True)
    assert np.abs(outs[0]["pcm"] - want).max() <= 1e-5


def test_mixed_layer_corpus(m, ref):
    """BASELINE.json configs[4] in small: one call over Layer I, II and III streams"""
    names = L12_CONFORMANCE + ["l3-si_huff", "l3-he_mode", "M2L3_compl24", "l3-compl"]
    datas = [load_vector(n) for n in names]*2
    rets, outs = m.mp3dec_decode_batch_cuda(datas, raw_frames=True)
    for d, r, o in zip(datas, rets, outs):
        want, _ = ref.decode_frames(d)
        assert r == 0 and len(o["pcm"]) == len(want)
        assert np.abs(o["pcm"].astype(np.int32) - want.astype(np.int32)).max() <= 1


def test_regrouped_lane_order_lists_every_granule_channel_once(m):
    # k_order.cu: the streams of a window of 32 are merged and sorted by (big_values / 2, part_23_length / 16), equal keys
    # by 256-slot stretch of the stream timeline; every real granule-channel must still be listed exactly once, and
    # copies of one stream must end up in whole warps of granule-channels of one size class
    data = [load_vector("l3-sin1k0db")]*64 + [load_vector("l3-compl"), load_vector("l3-he_mode"), load_vector("l1-fl1")]*3
    plan = m.BatchPlan(data, device=0)
    try:
        plan.run()
        plan.sync()
        order = plan.read_tap(8).astype(np.int64)
        assert len(np.unique(order)) == len(order)
        n_real = {}
        for d in set(data):
            one = m.BatchPlan([d], device=0)
            one.run()
            one.sync()
            n_real[d] = (len(one.read_tap(8)), int(one.stats["n_granule_channels"]))
            one.close()
        assert len(order) == sum(n_real[d][0] for d in data)
        # first window: 32 copies of one stream -> every warp's 32 lanes have the same two size fields, and the window's
        # part of the list holds every slot of every copy once
        per = n_real[data[0]][1]
        first = order[:32*n_real[data[0]][0]].reshape(-1, 32)
        assert first.max() < 32*per
        gcs = plan.read_tap(4)          # GcDesc bytes: part_23_length at 8, big_values at 10
        p23 = gcs[:, 8].astype(np.int64) | gcs[:, 9].astype(np.int64) << 8
        big = gcs[:, 10].astype(np.int64) | gcs[:, 11].astype(np.int64) << 8
        key = (big[first] >> 1)*1024 + (p23[first] >> 4)
        assert np.all(key == key[:, :1])
        assert np.all(np.diff(key[:, 0]) <= 0)          # longest first
        assert np.array_equal(np.sort(first.ravel()), np.sort(np.concatenate([c*per + np.sort(order[order < per]) for c in range(32)])))
    finally:
        plan.close()


def test_device_pow43_is_bit_exact_over_the_whole_range(m, ref):
    # L3_pow_43 (minimp3.h:716-740) on the device, every magnitude a 13-bit escape can produce; the device
    # divides without the compiler's slow-path call (l3_div_small), so all 8207 quotients are checked
    n = 8207
    got = np.zeros(n, np.float32)
    assert m.load().mp3dec_b200_pow43_device(got.ctypes.data, n) == 0
    ref.lib.ref_const_pow43.restype = C.c_float
    ref.lib.ref_const_pow43.argtypes = [C.c_int]
    want = np.array([ref.lib.ref_const_pow43(x) for x in range(n)], np.float32)
    bad = np.where(got.view(np.uint32) != want.view(np.uint32))[0]
    assert len(bad) == 0, bad[:10]


def test_f32_to_s16(m, ref_f32):
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-1.2, 1.2, 10007), np.array([0.5/32768, -0.5/32768, 1.5/32768, -1.5/32768, 1.0, -1.0])]).astype(np.float32)
    out = np.zeros(len(x), np.int16)
    m.load().mp3dec_f32_to_s16(x.ctypes.data, out.ctypes.data, len(x))
    want = np.zeros(len(x), np.int16)
    ref_f32.lib.mp3dec_f32_to_s16.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    ref_f32.lib.mp3dec_f32_to_s16(x.ctypes.data, want.ctypes.data, len(x))
    assert np.array_equal(out, want)


def test_fuzzed_batch_decodes_without_faults(m, ref):
    """junk with planted sync words, spliced / cut / duplicated streams, streams entered mid-frame, all in one batch
    with load_buf semantics: return codes and sample counts of the reference, no CUDA error, clean streams exact"""
    rng = np.random.default_rng(4242)
    base = load_vector("l3-he_mode")
    datas = [base]
    for t in range(96):
        kind = t % 4
        if kind == 0:
            n = int(rng.integers(1, 20000))
            b = bytearray(rng.integers(0, 256, n, dtype=np.uint8).tobytes())
            for _ in range(int(rng.integers(0, 40))):
                p = int(rng.integers(0, max(1, n - 4)))
                b[p:p + 2] = bytes([0xFF, int(rng.choice([0xFB, 0xFA, 0xF3, 0xE3, 0xFD, 0xFC]))])
        elif kind == 1:
            b = bytearray(base[:int(rng.integers(2000, len(base)))])
            for _ in range(int(rng.integers(1, 6))):
                p, n = int(rng.integers(0, len(b) - 200)), int(rng.integers(1, 200))
                b[p:p + n] = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        elif kind == 2:
            b = bytearray(base)
            for _ in range(int(rng.integers(1, 4))):
                p, n = int(rng.integers(0, len(b) - 500)), int(rng.integers(1, 500))
                if rng.random() < 0.5:
                    del b[p:p + n]
                else:
                    b[p:p] = b[p:p + n]
        else:
            b = bytearray(base[int(rng.integers(0, 3000)):])
        datas.append(bytes(b))
    datas.append(base)
    rets, outs = m.mp3dec_decode_batch_cuda(datas, allow_mono_stereo=True)   # the reference shim is built that way
    close = 0
    for i, (d, r, o) in enumerate(zip(datas, rets, outs)):
        ret, want, info = ref.load_buf(d)
        assert r == ret and len(o["pcm"]) == len(want), i
        if len(want):
            diff = np.abs(o["pcm"].astype(np.int32) - want.astype(np.int32))
            if i in (0, len(datas) - 1):
                assert diff.max() <= 1
            close += int((diff <= 1).mean() > 0.9)
        else:
            close += 1
    assert close >= len(datas)*0.8


@pytest.mark.parametrize("name", L3_CONFORMANCE)
def test_device_side_info_matches_host_parse(name, m):
    """the descriptors k_l3_side_info builds on the GPU from the raw side-info bytes are byte for byte
    those of the full host parse (tests/harness.py walks the stream with the host-side reader)."""
    import harness
    data = load_vector(name)
    want = np.ascontiguousarray(harness.walk(data, raw=True)["gcs"])
    p = m.BatchPlan([data], raw_frames=True)
    p.run()
    p.sync()
    desc = np.ascontiguousarray(p.read_tap(4))
    p.close()
    assert len(desc) >= len(want) and len(want) > 0
    assert desc[:len(want)].tobytes() == want.tobytes()


def _frame_loop(m, data, hop=None, rebuffer_at=None, chunk=None, parse_only_at=None):
    """mp3dec_init + mp3dec_decode_frame loop with the ways a caller can leave the straight path: move the rest of the
    stream to another buffer, feed at most `chunk` bytes per call, make one parse-only call (pcm = NULL)"""
    L = m.load()
    dec = m.Mp3Dec()
    L.mp3dec_init(C.byref(dec))
    info = m.FrameInfo()
    buf = np.frombuffer(data, dtype=np.uint8).copy()
    tmp = np.zeros(m.MINIMP3_MAX_SAMPLES_PER_FRAME, np.int16)
    pos, out, n, calls = 0, [], len(data), 0
    keep = [buf]
    while True:
        if rebuffer_at is not None and calls == rebuffer_at:
            buf = np.concatenate([np.zeros(pos, np.uint8), np.frombuffer(data, dtype=np.uint8)[pos:]])   # same bytes, new address
            keep.append(buf)
        avail = n - pos if chunk is None else min(n - pos, chunk)
        pcm_ptr = None if parse_only_at == calls else tmp.ctypes.data
        samples = L.mp3dec_decode_frame(C.byref(dec), buf.ctypes.data + pos, avail, pcm_ptr, C.byref(info))
        calls += 1
        if samples and pcm_ptr:
            out.append(tmp[:samples*info.channels].copy())
        pos += info.frame_bytes
        if not info.frame_bytes:
            break
    return np.concatenate(out) if out else np.zeros(0, np.int16)


@pytest.mark.parametrize("name", ["l3-compl", "l3-he_mode", "l3-sin1k0db", "M2L3_bitrate_22_all", "l2-fl12"])
def test_frame_loop_look_ahead_and_its_exits(name, m, ref, monkeypatch):
    """the plain frame loop is answered from look-ahead windows; leaving the predicted path (new buffer, short reads,
    tiny windows) replays the decoder state and goes on through the one-frame path: always the reference's stream"""
    data = load_vector(name)
    want, _ = ref.decode_frames(data)

    def check(pcm):
        assert len(pcm) == len(want) and np.abs(pcm.astype(np.int32) - want.astype(np.int32)).max() <= 1

    check(_frame_loop(m, data))                                   # straight through one window
    check(_frame_loop(m, data, rebuffer_at=37))                   # the stream moves to another buffer mid-way
    check(_frame_loop(m, data, rebuffer_at=1))
    check(_frame_loop(m, data, chunk=6000))                       # never enough bytes for a window: one-frame path
    monkeypatch.setenv("MP3B_DEBUG_WINDOW_FRAMES", "9")           # many window boundaries
    check(_frame_loop(m, data))
    check(_frame_loop(m, data, rebuffer_at=23))


def test_frame_loop_on_damaged_streams_call_by_call(m, ref):
    """spliced / cut / duplicated streams through the frame loop (look-ahead windows engaged): every call reports what
    the reference's call reports (samples, frame_bytes), sample counts match"""
    rng = np.random.default_rng(777)
    base = load_vector("l3-sin1k0db")
    for t in range(12):
        b = bytearray(base)
        kind = t % 3
        if kind == 0:
            for _ in range(int(rng.integers(1, 6))):
                p, n = int(rng.integers(0, len(b) - 200)), int(rng.integers(1, 200))
                b[p:p + n] = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        elif kind == 1:
            for _ in range(int(rng.integers(1, 4))):
                p, n = int(rng.integers(0, len(b) - 500)), int(rng.integers(1, 500))
                if rng.random() < 0.5:
                    del b[p:p + n]
                else:
                    b[p:p] = b[p:p + n]
        else:
            b = b[int(rng.integers(1, 3000)):int(rng.integers(60000, len(b)))]
        data = bytes(b)
        want_pcm, want_frames = ref.decode_frames(data)
        pcm, frames = m.decode_frames(data)
        assert [(f["samples"], f["frame_bytes"]) for f in frames] == [(f.samples, f.frame_bytes) for f in want_frames], t
        assert len(pcm) == len(want_pcm)


def test_frame_loop_on_a_long_stream(m, ref):
    """12 720 frames (> 3 default windows of 4096, input views shorter than the buffer): frame loop and streaming read"""
    data = load_vector("l3-sin1k0db")*40
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
        n = L.mp3dec_ex_read(C.byref(dec), out.ctypes.data + 2*got, 100000)
        got += n
        if n != 100000:
            break
    L.mp3dec_ex_close(C.byref(dec))
    want_lb = ref.load_buf(data)[1]
    assert got == len(want_lb) and np.abs(out[:got].astype(np.int32) - want_lb.astype(np.int32)).max() <= 1


def test_concurrent_frame_loops(m, ref):
    """several decoders walked from different host threads at the same time (each with its own look-ahead window)"""
    import threading
    names = ["l3-compl", "l3-he_mode", "l3-sin1k0db", "M2L3_bitrate_22_all", "l2-fl12", "l3-si_huff", "l3-he_44khz", "l1-fl1"]
    datas = [load_vector(n) for n in names]
    results = [None]*len(names)

    def work(i):
        results[i] = m.decode_frames(datas[i])[0]

    for rep in range(2):
        threads = [threading.Thread(target=work, args=(i,)) for i in range(len(names))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for i, d in enumerate(datas):
            want, _ = ref.decode_frames(d)
            assert len(results[i]) == len(want) and np.abs(results[i].astype(np.int32) - want.astype(np.int32)).max() <= 1, names[i]


def test_more_decoders_than_live_windows(m, ref):
    """20 decoders advanced round-robin, a few frames at a time: more than the 16 windows kept alive, so windows are
    retired to saved decoder states and decoding continues on the one-frame path -- every stream still exact"""
    L = m.load()
    data = load_vector("l3-he_mode")
    want, _ = ref.decode_frames(data)
    buf = np.frombuffer(data, dtype=np.uint8)
    n_dec = 20
    decs = [m.Mp3Dec() for _ in range(n_dec)]
    for d in decs:
        L.mp3dec_init(C.byref(d))
    pos = [0]*n_dec
    outs = [[] for _ in range(n_dec)]
    done = [False]*n_dec
    info = m.FrameInfo()
    tmp = np.zeros(m.MINIMP3_MAX_SAMPLES_PER_FRAME, np.int16)
    while not all(done):
        for i in range(n_dec):
            for _ in range(3):
                if done[i]:
                    break
                s = L.mp3dec_decode_frame(C.byref(decs[i]), buf.ctypes.data + pos[i], len(data) - pos[i], tmp.ctypes.data, C.byref(info))
                if s:
                    outs[i].append(tmp[:s*info.channels].copy())
                pos[i] += info.frame_bytes
                if not info.frame_bytes:
                    done[i] = True
    for i in range(n_dec):
        pcm = np.concatenate(outs[i])
        assert len(pcm) == len(want) and np.abs(pcm.astype(np.int32) - want.astype(np.int32)).max() <= 1, i
