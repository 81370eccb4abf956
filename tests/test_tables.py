"""Generated constant tables (tools/gen_tables.py) against the compiled reference and against themselves."""
import os
import re

import numpy as np
import pytest

import conftest
import harness


def parse_array(text, name):
    m = re.search(r"%s\[[^\]]*\] = \{(.*?)\};" % name, text, re.S)
    body = m.group(1).replace("\n", " ")
    return [x.strip() for x in body.split(",") if x.strip()]


@pytest.fixture(scope="module")
def tables_text():
    return open(os.path.join(conftest.ROOT, "minimp3_b200", "csrc", "mp3_tables.h")).read()


def test_oracle_and_product_tables_are_the_same_file(tables_text):
    assert tables_text == open(os.path.join(conftest.ROOT, "oracle", "mp3_tables.h")).read()


def test_pow43_matches_reference(ref):
    # L3_pow_43 (minimp3.h:721-740): table below 129, polynomial above; bit-exact over the whole range
    L = harness.lib()
    for x in list(range(0, 2000)) + list(range(2000, 8207, 7)) + [8206]:
        a = np.float32(L.h_pow43(x))
        b = np.float32(ref.lib.ref_const_pow43(x))
        assert a.view(np.uint32) == b.view(np.uint32), x


def test_ldexp_q2_matches_reference(ref):
    # L3_ldexp_q2 (minimp3.h:642-652)
    L = harness.lib()
    for y in (1.0, 2048.0, 0.7071, 3.1e-5):
        for e in list(range(0, 300)) + [400, 511, 1020]:
            a = np.float32(L.h_ldexp_q2(y, e))
            b = np.float32(ref.lib.ref_const_ldexp_q2(y, e))
            assert a.view(np.uint32) == b.view(np.uint32), (y, e)


def test_huffman_lut_decodes_every_iso_code(tables_text):
    """Walk the multi-level LUT with every code word of the raw ISO list: must return (x, y) and the code length."""
    lut = [int(x, 16) for x in parse_array(tables_text, "MP3T_HUFF_LUT")]
    base = [int(x) for x in parse_array(tables_text, "MP3T_HUFF_BASE")]
    rootw = [int(x) for x in parse_array(tables_text, "MP3T_HUFF_ROOTW")]
    cbase = [int(x) for x in parse_array(tables_text, "MP3T_HUFF_CODES_BASE")]
    ccount = [int(x) for x in parse_array(tables_text, "MP3T_HUFF_CODES_COUNT")]
    raw = re.findall(r"\{(\d+),0x([0-9a-f]+),(\d+),(\d+)\}", tables_text)
    codes = [(int(l), int(c, 16), int(x), int(y)) for l, c, x, y in raw]
    seen_tables = 0
    for t in range(32):
        if not ccount[t]:
            assert rootw[t] == 0
            continue
        seen_tables += 1
        kraft = 0.0
        for (ln, code, x, y) in codes[cbase[t]:cbase[t] + ccount[t]]:
            kraft += 2.0**-ln
            bits = code << (32 - ln)          # left-aligned 32-bit window, zero padded
            pos, w = 0, rootw[t]
            e = lut[base[t] + ((bits >> (32 - w)) & ((1 << w) - 1))]
            while e & 0x8000:
                pos += w
                w = (e >> 11) & 15
                idx = (bits >> (32 - pos - w)) & ((1 << w) - 1) if pos + w <= 32 else 0
                e = lut[base[t] + (e & 0x7FF) + idx]
            pos += e >> 8
            assert (pos, (e >> 4) & 15, e & 15) == (ln, x, y), (t, ln, hex(code))
        assert abs(kraft - 1.0) < 1e-12, t      # complete prefix code: no undecodable bit pattern
    assert seen_tables == 29


def test_sfb_tables_cover_576_lines(tables_text):
    for name, width in (("MP3T_SFB_LONG", 23), ("MP3T_SFB_SHORT", 40), ("MP3T_SFB_MIXED", 40)):
        vals = [int(x) for x in parse_array(tables_text, name)]
        for row in range(8):
            r = vals[row*width:(row + 1)*width]
            assert sum(r) == 576 and 0 in r and all(v % 2 == 0 for v in r)


def test_window_table_properties(tables_text):
    dw = np.array([int(x) for x in parse_array(tables_text, "MP3T_DWIN")])
    assert dw[0] == 0 and dw[256] == 75038 and dw[64] == 213
    # ISO D[i] symmetry |D[512 - i]| == |D[i]|, except the taps n = 64m + 16 that multiply V[16] == 0
    # (unobservable through the filterbank, stored as 0)
    n = np.arange(1, 256)
    keep = (n % 32) != 16
    assert np.array_equal(np.abs(dw[n[keep]]), np.abs(dw[512 - n[keep]]))
    assert np.all(dw[16::64] == 0)
