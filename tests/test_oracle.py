"""Pinning the oracle (oracle/mp3_oracle.c, the plain-C restatement) against the reference's own goldens:
sample counts from the golden manifest (produced by the unmodified reference), PSNR >= 96 dB against the ISO
conformance PCM bundled with the reference (minimp3_test.c:417-427), and -- when the compiled reference is
available -- its stage taps: integers and requantised lines bit-exact, PCM within +-1 LSB."""
import numpy as np
import pytest

from conftest import L12_CONFORMANCE, L12_REGRESSION, L3_CONFORMANCE, load_iso_pcm, load_vector, psnr
from oracle import port


@pytest.mark.parametrize("name", L3_CONFORMANCE)
def test_port_against_golden_manifest_and_iso_pcm(name, manifest):
    o = port.decode_stream(load_vector(name))
    m = manifest[name]
    assert o["ret"] == 0
    assert len(o["pcm"]) == m["samples"]                     # what the unmodified reference produced here
    iso = load_iso_pcm(name)
    assert len(o["pcm"]) in (len(iso), len(iso) + 1152, len(iso) + 2304)
    md, ps = psnr(o["pcm"], iso)
    assert ps >= 96.0 and md <= 2
    assert ps >= m["ref_psnr"] - 1.0


@pytest.mark.parametrize("name", L3_CONFORMANCE)
def test_port_against_reference_taps(name, ref):
    data = load_vector(name)
    o = port.decode_stream(data, taps=True)
    t = ref.tap_stream(data)
    assert o["n_gc"] == len(t["rec"])
    assert np.array_equal(o["q"], t["q"])
    assert np.array_equal(o["xr"].view(np.uint32), t["xr"].view(np.uint32))
    for i, r in enumerate(t["rec"]):
        n = r.n_long_sfb + r.n_short_sfb
        assert np.array_equal(o["scf"][i, :n].view(np.uint32), t["scf"][i, :n].view(np.uint32))
    scale = max(1.0, float(np.abs(t["sub"]).max()))
    assert np.abs(o["sub"] - t["sub"]).max() <= 2e-6*scale
    d = np.abs(o["pcm"].astype(np.int32) - t["pcm"].astype(np.int32))
    assert d.max() <= 1 and (d != 0).mean() < 0.01


def test_port_float_output(ref_f32):
    data = load_vector("l3-si_block")
    want, _ = ref_f32.decode_frames(data)
    o = port.decode_stream(data, float_output=True)
    assert np.abs(o["pcm"] - want).max() <= 1e-5


@pytest.mark.parametrize("name", L12_CONFORMANCE)
def test_port_layer12_against_golden_manifest_and_iso_pcm(name, manifest):
    """Layer I / II restated from the ISO allocation tables: sample count of the unmodified reference, PSNR against
    the ISO conformance PCM (minimp3_test.c:417-427)."""
    o = port.decode_stream(load_vector(name))
    m = manifest[name]
    assert o["ret"] == 0 and len(o["pcm"]) == m["samples"]
    md, ps = psnr(o["pcm"], load_iso_pcm(name))
    assert ps >= 96.0 and md <= 1


@pytest.mark.parametrize("name", L12_CONFORMANCE + L12_REGRESSION)
def test_port_layer12_against_reference_pcm(name, ref):
    """every Layer I / II vector incl. the malformed ILL* ones: same frames kept and dropped, PCM within +-1 LSB"""
    data = load_vector(name)
    want, _ = ref.decode_frames(data)
    o = port.decode_stream(data)
    assert o["ret"] == 0 and len(o["pcm"]) == len(want)
    if len(want):
        assert np.abs(o["pcm"].astype(np.int32) - want.astype(np.int32)).max() <= 1
