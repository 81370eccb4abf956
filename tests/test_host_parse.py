"""Host side of the hot path (frame sync, side info, bit reservoir, load_buf bookkeeping, tile planning)
against the compiled reference -- integer work, so everything must match exactly."""
import numpy as np
import pytest

import conftest
import harness
from conftest import L3_CONFORMANCE, L12_CONFORMANCE, L12_REGRESSION, load_vector, real_gc_index

NONSTD = ["l3-nonstandard-apetag", "l3-nonstandard-big-iscf", "l3-nonstandard-compl-sideinfo-bigvalues",
          "l3-nonstandard-compl-sideinfo-blocktype", "l3-nonstandard-compl-sideinfo-size", "l3-nonstandard-he_44_48khz",
          "l3-nonstandard-id3v1-apetag", "l3-nonstandard-id3v1", "l3-nonstandard-id3v2-only", "l3-nonstandard-id3v2",
          "l3-nonstandard-sideinfo-size", "l3-nonstandard-sin1k0db_lame_vbrtag", "l3-nonstandard-small",
          "l3-nonstandard-vbrtag-corrupted", "l3-nonstandard-vbrtag-empty", "l3-nonstandard-vbrtag-noframes",
          "l3-nonstandard-vbrtag-only", "l3-nonstandard-vbrtag-oob-read"]


@pytest.mark.parametrize("name", L3_CONFORMANCE + NONSTD + ["MIPSTest.mp3"] + L12_CONFORMANCE + L12_REGRESSION)
def test_raw_frame_walk_matches_reference(name, ref):
    """frames, decoded frames and sample count of the plain mp3dec_decode_frame loop (minimp3.h:1713-1806)"""
    data = load_vector(name)
    pcm, frames = ref.decode_frames(data)
    w = harness.walk(data, raw=True)
    assert w["info"]["frames"] == len(frames)
    assert w["info"]["frames_decoded"] == sum(1 for f in frames if f.samples)
    assert w["info"]["samples_decoded"] == len(pcm)
    assert w["info"]["ret"] == 0


@pytest.mark.parametrize("name", L3_CONFORMANCE + NONSTD + L12_CONFORMANCE + L12_REGRESSION)
def test_load_buf_bookkeeping_matches_reference(name, ref, manifest):
    """return code, delivered samples, channels/hz/layer/avg_bitrate of mp3dec_load_buf (minimp3_ex.h:308-525).
    The reference test binary is built with MINIMP3_ALLOW_MONO_STEREO_TRANSITION, so we compare with it on."""
    data = load_vector(name)
    ret, pcm, info = ref.load_buf(data)
    w = harness.walk(data, raw=False, allow=True)
    assert w["info"]["ret"] == ret
    assert w["info"]["out_samples"] == len(pcm)
    for k in ("channels", "hz", "layer", "avg_bitrate_kbps"):
        assert w["info"][k] == info[k], k
    m = manifest[name]
    assert m["load_ret"] == ret and m["load_samples"] == len(pcm)


def test_mono_stereo_transition_is_an_error_by_default():
    """without MINIMP3_ALLOW_MONO_STEREO_TRANSITION a channel change ends the file with MP3D_E_DECODE (minimp3_ex.h:484-492)"""
    w = harness.walk(load_vector("l3-he_mode"), raw=False, allow=False)
    assert w["info"]["ret"] == -5
    w = harness.walk(load_vector("l3-he_mode"), raw=False, allow=True)
    assert w["info"]["ret"] == 0 and w["info"]["channels"] == 0


@pytest.mark.parametrize("name", L3_CONFORMANCE)
def test_descriptors_match_reference_side_info(name, ref):
    """every GcDesc field against the reference's L3_gr_info_t, and the resolved reservoir bit offset against
    the bytes the reference actually decoded from (L3_restore_reservoir, minimp3.h:1228-1236)"""
    data = load_vector(name)
    t = ref.tap_stream(data)
    w = harness.walk(data, raw=True)
    idx = real_gc_index(w)
    assert len(idx) == len(t["rec"])
    g = w["gcs"][idx]
    for i, r in enumerate(t["rec"]):
        d = g[i]
        assert d["part_23_length"] == r.part_23_length and d["big_values"] == r.big_values
        assert d["scalefac_compress"] == r.scalefac_compress and d["global_gain"] == r.global_gain
        assert (d["flags0"] & 3) == r.block_type and bool(d["flags0"] & 4) == bool(r.mixed_block_flag)
        assert bool(d["flags0"] & 8) == bool(r.preflag) and bool(d["flags0"] & 16) == bool(r.scalefac_scale)
        assert bool(d["flags0"] & 32) == bool(r.count1_table)
        assert list(d["table_select"]) == list(r.table_select)
        if r.block_type == 0:
            assert list(d["region_count"]) == list(r.region_count)
        else:
            assert list(d["region_count"][:2]) == list(r.region_count[:2])
        assert list(d["subblock_gain"]) == list(r.subblock_gain) or r.block_type == 0
        assert d["scfsi"] == r.scfsi and d["n_long_sfb"] == r.n_long_sfb and d["n_short_sfb"] == r.n_short_sfb
        assert d["sr_idx"] == r.sr_idx
        assert bool(d["flags1"] & 2) == bool(r.ch) and bool(d["flags1"] & 4) == (r.nch == 2) and bool(d["flags1"] & 8) == bool(r.gr)


def test_tiles_partition_the_granule_list():
    for name in ("l3-he_mode", "l3-sin1k0db", "l3-compl", "M2L3_bitrate_16_all"):
        w = harness.walk(load_vector(name), raw=True)
        n_gr = 0
        pcm = 0
        for t in w["tiles"]:
            assert 1 <= t["n_gr"] <= (16 if t["nch"] == 1 else 8) and t["nch"] in (1, 2)   # mono tiles hold twice the granules
            assert t["pcm_off"] == pcm
            assert t["gc_first"] == w["gr_gc"][n_gr]
            for k in range(int(t["n_gr"])):
                assert w["gr_nch"][n_gr + k] == t["nch"]
                assert int(w["gr_gc"][n_gr + k]) == int(t["gc_first"]) + k*int(t["nch"])
            # halo = the two previous granule-channels of each channel since the last reset
            for ch in range(t["nch"]):
                prev = [int(w["gr_gc"][j]) + ch for j in range(n_gr) if w["gr_nch"][j] > ch]
                want = [prev[-1] if len(prev) >= 1 else -1, prev[-2] if len(prev) >= 2 else -1]
                assert list(t["halo"][ch]) == want
            n_gr += int(t["n_gr"])
            pcm += 576*int(t["nch"])*int(t["n_gr"])
        assert n_gr == len(w["gr_gc"]) and pcm == w["info"]["samples_decoded"]
        # stereo pairs sit on even descriptor indices
        assert all(int(g) % 2 == 0 for g, n in zip(w["gr_gc"], w["gr_nch"]) if n == 2)


@pytest.mark.parametrize("name", ["l3-compl", "l3-nonstandard-id3v2-only", "l3-nonstandard-small", "l3-nonstandard-apetag",
                                  "l1-fl1", "l2-fl10", "l3-nonstandard-vbrtag-only"])
def test_detect_buf_matches_reference(name, ref):
    data = load_vector(name)
    buf = np.frombuffer(data, dtype=np.uint8)
    assert harness.lib().h_detect_buf(buf.ctypes.data, len(data)) == ref.detect_buf(data)


def test_empty_and_garbage_inputs():
    assert harness.walk(b"", raw=False)["info"]["out_samples"] == 0
    assert harness.walk(b"\x00"*1000, raw=True)["info"]["samples_decoded"] == 0
    rng = np.random.default_rng(0)
    junk = rng.integers(0, 256, 20000, dtype=np.uint8).tobytes()
    w = harness.walk(junk, raw=True)
    assert w["info"]["ret"] in (0, -6)


def test_layer12_streams_are_planned():
    """Layer I/II frames become 384-sample blocks with their own synthesis tiles"""
    w = harness.walk(load_vector("l2-fl10"), raw=True)
    assert w["info"]["ret"] == 0 and w["info"]["layer"] == 2 and w["info"]["n_gc"] == 0
    assert w["info"]["samples_decoded"] == 112896


@pytest.mark.parametrize("name", ["l3-compl", "l3-he_mode", "l3-sin1k0db", "M2L3_bitrate_16_all", "l3-he_free", "l2-fl10", "l1-fl1",
                                  "l2-nonstandard-free_format"])
def test_corrupted_streams_walk_like_the_reference(name, ref):
    """robustness: random byte damage (resyncs, bad side info, reservoir underruns, overrunning Layer II frames) must
    give the same frame boundaries and sample counts as the reference's own parser"""
    data = bytearray(load_vector(name))
    import zlib
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    for trial in range(24):
        d = bytearray(data)
        for _ in range(12):
            p = int(rng.integers(0, len(d)))
            d[p] = int(rng.integers(0, 256))
        if trial % 2:
            cut = int(rng.integers(len(d)//2, len(d)))
            d = d[:cut]
        d = bytes(d)
        pcm, frames = ref.decode_frames(d)
        w = harness.walk(d, raw=True)
        assert w["info"]["frames"] == len(frames), trial
        assert w["info"]["frames_decoded"] == sum(1 for f in frames if f.samples), trial
        assert w["info"]["samples_decoded"] == len(pcm), trial


def test_fuzzed_streams_walk_like_the_reference(ref):
    """Random junk with planted sync words, spliced / cut / duplicated real streams, streams entered mid-frame:
    the walker keeps and drops the same frames as mp3dec_load_buf (return code, sample count, format)."""
    rng = np.random.default_rng(12345)
    base = load_vector("l3-compl")
    for t in range(160):
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
        data = bytes(b)
        ret, pcm, info = ref.load_buf(data)
        w = harness.walk(data)
        wi = w["info"]
        assert (wi["ret"], wi["out_samples"]) == (ret, len(pcm)), (t, kind)
        if len(pcm):
            assert (wi["hz"], wi["channels"]) == (info["hz"], info["channels"]), (t, kind)
        # the batch path's walker (lite side info + frame records) keeps the same frames and builds the same records
        import ctypes as C
        L = harness.lib()
        L.h_walk_frame_recs.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        n = len(w["gcs"])
        gcs = np.zeros(max(1, n), harness.GCDESC_DTYPE)
        sc = np.zeros(16, np.int64)
        buf = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, np.uint8)
        assert L.h_walk_frame_recs(buf.ctypes.data, len(data), 0, 0, gcs.ctypes.data, n, None, 0, sc.ctypes.data) >= 0
        assert [int(v) for v in sc[:15]] == list(wi.values()), (t, kind)
        assert gcs[:n].tobytes() == np.ascontiguousarray(w["gcs"]).tobytes(), (t, kind)


@pytest.mark.parametrize("name", L3_CONFORMANCE + ["l3-nonstandard-big-iscf", "l3-nonstandard-compl-sideinfo-bigvalues",
                                                   "l3-nonstandard-compl-sideinfo-blocktype", "l3-nonstandard-sideinfo-size"])
def test_frame_records_give_the_same_descriptors(name):
    """in the batch path the walker reads only a third of the side info (read_side_info_lite_t) and hands
    the raw bytes to the GPU; the descriptors k_l3_side_info builds from them (same code, run on the host here), the
    bookkeeping and the lane order must be exactly those of the full host parse."""
    import ctypes as C
    data = load_vector(name)
    for raw in (0, 1):
        w = harness.walk(data, raw=bool(raw), allow=True)
        L = harness.lib()
        L.h_walk_frame_recs.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        assert L.h_sizeof_framerec() == 64
        buf = np.frombuffer(data, dtype=np.uint8)
        n = len(w["gcs"])
        gcs = np.zeros(max(1, n), harness.GCDESC_DTYPE)
        order = np.zeros(max(1, n), np.uint32)
        sc = np.zeros(16, np.int64)
        n_order = L.h_walk_frame_recs(buf.ctypes.data, len(data), raw, 1, gcs.ctypes.data, n, order.ctypes.data, n, sc.ctypes.data)
        assert n_order >= 0
        keys = ["ret", "samples_decoded", "out_skip", "out_samples", "channels", "hz", "layer", "avg_bitrate_kbps",
                "main_data_bytes", "frames", "frames_decoded", "n_gc", "n_tiles", "n_granules", "n_istereo"]
        assert {k: int(sc[i]) for i, k in enumerate(keys)} == w["info"]
        assert gcs[:n].tobytes() == np.ascontiguousarray(w["gcs"]).tobytes()
