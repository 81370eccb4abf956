"""The __host__ __device__ arithmetic the CUDA lanes execute (l3_entropy.cuh, l3_math.cuh), stepped on the CPU
through tests/cpu_harness.cpp and compared with the reference's stage taps:
integer spectra, scalefactors and requantised lines bit-exact; PCM within +-1 LSB."""
import numpy as np
import pytest

import harness
from conftest import L3_CONFORMANCE, load_vector, real_gc_index


@pytest.mark.parametrize("name", L3_CONFORMANCE + ["MIPSTest.mp3"])
def test_entropy_decode_is_bit_exact(name, ref):
    data = load_vector(name)
    t = ref.tap_stream(data)
    w = harness.walk(data, raw=True)
    idx = real_gc_index(w)
    e = harness.entropy(w["md"], w["gcs"])
    assert np.array_equal(e["q"][idx], t["q"])                                      # Huffman-decoded integers
    assert np.array_equal(e["xr"][idx].view(np.uint32), t["xr"].view(np.uint32))    # L3_pow_43 * scf, bitwise
    for i, r in enumerate(t["rec"]):
        n = r.n_long_sfb + r.n_short_sfb
        assert np.array_equal(e["scf"][idx[i], :n].view(np.uint32), t["scf"][i, :n].view(np.uint32))
    # decoded extent covers every non-zero line
    for i, g in enumerate(idx):
        nz = np.nonzero(t["q"][i])[0]
        assert len(nz) == 0 or nz[-1] < e["extent"][g]


@pytest.mark.parametrize("name", ["l3-he_mode", "l3-test45", "l3-test46"])
def test_intensity_positions_match(name, ref):
    data = load_vector(name)
    t = ref.tap_stream(data)
    w = harness.walk(data, raw=True)
    idx = real_gc_index(w)
    e = harness.entropy(w["md"], w["gcs"])
    n_checked = 0
    for i, r in enumerate(t["rec"]):
        if r.ch == 1 and (r.hdr & 0x10):
            n = r.n_long_sfb + r.n_short_sfb - (3 if r.n_short_sfb else 1)
            # the reference overwrites the top band(s) inside L3_intensity_stereo; compare what was read
            assert np.array_equal(e["ist"][idx[i], :n], t["ist_pos"][i, :n]), i
            n_checked += 1
    assert n_checked > 0


@pytest.mark.parametrize("name", L3_CONFORMANCE)
def test_transform_chain_within_one_lsb(name, ref):
    """reorder + antialias + IMDCT + DCT-II + window on the reference's post-stereo spectra"""
    data = load_vector(name)
    t = ref.tap_stream(data)
    w = harness.walk(data, raw=True)
    idx = real_gc_index(w)
    xs = np.zeros((len(w["gcs"]), 576), np.float32)
    xs[idx] = t["xr_stereo"]
    r = harness.transform(xs, w)
    scale = max(1.0, float(np.abs(t["sub"]).max()))
    assert np.abs(r["xr_aa"][idx] - t["xr_aa"]).max() <= 1e-6*max(1.0, float(np.abs(t["xr_aa"]).max()))
    assert np.abs(r["sub"][idx] - t["sub"]).max() <= 2e-6*scale
    d = np.abs(r["pcm16"].astype(np.int32) - t["pcm"].astype(np.int32))
    assert d.max() <= 1
    assert (d != 0).mean() < 0.01


def test_dct32_matches_closed_form():
    rng = np.random.default_rng(7)
    x = rng.uniform(-1, 1, 32).astype(np.float32)
    y = x.copy()
    harness.lib().h_dct32(y.ctypes.data)
    n = np.arange(32)
    want = np.array([(x.astype(np.float64)*np.cos(np.pi*(2*n + 1)*k/64)).sum() for k in range(32)])
    assert np.abs(y - want).max() < 2e-5
