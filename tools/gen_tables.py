#!/usr/bin/env python3
"""Generate the constant tables of the Layer III hot path.

Everything is produced from ISO/IEC 11172-3 / 13818-3 facts and closed forms, NOT by
copying the reference's arrays:

* scalefactor-band partitions   -> ISO band-width tables typed in below, mixed-block tables
                                   derived from them by the ISO rule
* |q|^(4/3) table               -> closed form, rounded to the 6 decimals the format uses
* Huffman code tables (B.7)     -> parsed from the ISO text table (huffopt/HUFFCODE in the
                                   reference checkout, build container only) and re-packed into a
                                   GPU-shaped multi-level LUT (9-bit root, <=8-bit sub-tables)
                                   plus a plain (len, code, x, y) list for the oracle port
* synthesis window D[i]*65536   -> measured as integer impulse responses of the compiled
                                   reference (oracle/_ref) and checked for integrality and the
                                   ISO symmetry, then stored as a flat 512-entry table
* IMDCT / DCT / antialias twiddles -> closed forms in float32

When oracle/_ref is available every table is cross-checked against the compiled reference
(see also tests/test_tables.py).  Output: minimp3_b200/csrc/mp3_tables.h and
oracle/mp3_tables.h (same content; the oracle keeps its own copy so that it never
includes product code).
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("MINIMP3_REFERENCE", "/root/reference")

# --------------------------------------------------------------------------- sfb partitions
# ISO 11172-3 table B.8 / 13818-3 / MPEG-2.5 extension: widths of the scalefactor bands.
# Row order = the decoder's sample-rate class:
#   0: 11025/12000 (MPEG-2.5)  1: 8000 (MPEG-2.5)  2: 22050  3: 24000  4: 16000 (MPEG-2)
#   5: 44100  6: 48000  7: 32000 (MPEG-1)
LONG_W = {
    "lsf22": [6, 6, 6, 6, 6, 6, 8, 10, 12, 14, 16, 20, 24, 28, 32, 38, 46, 52, 60, 68, 58, 54],
    "lsf24": [6, 6, 6, 6, 6, 6, 8, 10, 12, 14, 16, 18, 22, 26, 32, 38, 46, 54, 62, 70, 76, 36],
    "m25_8": [12, 12, 12, 12, 12, 12, 16, 20, 24, 28, 32, 40, 48, 56, 64, 76, 90, 2, 2, 2, 2, 2],
    "m1_44": [4, 4, 4, 4, 4, 4, 6, 6, 8, 8, 10, 12, 16, 20, 24, 28, 34, 42, 50, 54, 76, 158],
    "m1_48": [4, 4, 4, 4, 4, 4, 6, 6, 6, 8, 10, 12, 16, 18, 22, 28, 34, 40, 46, 54, 54, 192],
    "m1_32": [4, 4, 4, 4, 4, 4, 6, 6, 8, 10, 12, 16, 20, 24, 30, 38, 46, 56, 68, 84, 102, 26],
}
SHORT_W = {
    "m25_11": [4, 4, 4, 6, 8, 10, 12, 14, 18, 24, 30, 40, 18],
    "m25_8": [8, 8, 8, 12, 16, 20, 24, 28, 36, 2, 2, 2, 26],
    "lsf22": [4, 4, 4, 6, 6, 8, 10, 14, 18, 26, 32, 42, 18],
    "lsf24": [4, 4, 4, 6, 8, 10, 12, 14, 18, 24, 32, 44, 12],
    "lsf16": [4, 4, 4, 6, 8, 10, 12, 14, 18, 24, 30, 40, 18],
    "m1_44": [4, 4, 4, 4, 6, 8, 10, 12, 14, 18, 22, 30, 56],
    "m1_48": [4, 4, 4, 4, 6, 6, 10, 12, 14, 16, 20, 26, 66],
    "m1_32": [4, 4, 4, 4, 6, 8, 12, 16, 20, 26, 34, 42, 12],
}
ROW_LONG = ["lsf22", "m25_8", "lsf22", "lsf24", "lsf22", "m1_44", "m1_48", "m1_32"]
ROW_SHORT = ["m25_11", "m25_8", "lsf22", "lsf24", "lsf16", "m1_44", "m1_48", "m1_32"]


def sfb_tables():
    long_t, short_t, mixed_t = [], [], []
    for row in range(8):
        lw = LONG_W[ROW_LONG[row]]
        sw = SHORT_W[ROW_SHORT[row]]
        assert sum(lw) == 576 and sum(sw) == 192
        long_t.append(lw + [0])
        short_t.append([w for w in sw for _ in range(3)] + [0])
        # mixed blocks: the first 36 lines (two subbands) keep the long partition, the rest uses the
        # short partition from the band that starts at line 12 of each window (12*3 = 36 lines).
        m, acc = [], 0
        for w in lw:
            if acc >= 36:
                break
            m.append(w)
            acc += w
        assert acc == 36
        acc, k = 0, 0
        while acc < 12:            # skip short bands covering lines < 12 of each window
            acc += sw[k]
            k += 1
        if acc > 12:               # 8 kHz: a short band straddles the boundary -> keep its upper part
            m += [acc - 12]*3
        m += [w for w in sw[k:] for _ in range(3)]
        assert sum(m) == 576, (row, sum(m))
        mixed_t.append(m + [0]*(40 - len(m)))
    return long_t, short_t, mixed_t


# --------------------------------------------------------------------------- pow43
def pow43_table():
    # the format's requantiser uses |q|^(4/3) rounded to 6 decimals (float32)
    return [np.float32(float("%.6f" % (i**(4.0/3.0)))) for i in range(129 + 16)]


# --------------------------------------------------------------------------- Huffman
def parse_iso_huffman(path):
    tabs, cur = {}, None
    for line in open(path):
        line = line.strip()
        if not line or line.startswith("#"):
            continue
        if line.startswith(".table"):
            p = line.split()
            cur = int(p[1])
            tabs[cur] = dict(linbits=int(p[4]), codes=[], ref=None)
        elif line.startswith(".reference"):
            tabs[cur]["ref"] = int(line.split()[1])
        elif line.startswith(".end"):
            break
        else:
            p = line.split()
            if len(p) == 3:
                p = ["0"] + p
            assert len(p[3]) == int(p[2])
            tabs[cur]["codes"].append((int(p[0]), int(p[1]), int(p[2]), int(p[3], 2)))
    return tabs


ROOT_BITS, SUB_BITS = 9, 8


def build_lut(codes, root_bits=ROOT_BITS, sub_bits=SUB_BITS):
    """Multi-level LUT.  Entry (uint16): leaf  = len<<8 | x<<4 | y   (bit15 clear)
                                         link  = 0x8000 | width<<11 | offset (relative to table base)"""
    out = []

    def rec(items, w):
        base = len(out)
        out.extend([None]*(1 << w))
        groups = {}
        for (x, y, r, c) in items:
            if r <= w:
                lo = c << (w - r)
                for k in range(1 << (w - r)):
                    assert out[base + lo + k] is None
                    out[base + lo + k] = (r << 8) | (x << 4) | y
            else:
                groups.setdefault(c >> (r - w), []).append((x, y, r - w, c & ((1 << (r - w)) - 1)))
        for pre, g in sorted(groups.items()):
            w2 = min(max(i[2] for i in g), sub_bits)
            sub = rec(g, w2)
            assert sub < 2048 and out[base + pre] is None
            out[base + pre] = 0x8000 | (w2 << 11) | sub
        return base

    w0 = min(max(c[2] for c in codes), root_bits)
    rec(list(codes), w0)
    assert all(e is not None for e in out), "code is not complete"
    return w0, out


LINBITS = [0]*16 + [1, 2, 3, 4, 6, 8, 10, 13, 4, 5, 6, 7, 8, 9, 11, 13]


def huffman_tables(path):
    tabs = parse_iso_huffman(path)
    for t in range(34):
        if t < 32:
            assert tabs[t]["linbits"] == LINBITS[t], t
    lut, base, rootw = [], [0]*32, [0]*32
    raw, raw_base, raw_count = [], [0]*32, [0]*32
    built = {}
    for t in range(32):
        src = tabs[t]["ref"] if tabs[t]["ref"] is not None else t
        codes = tabs[src]["codes"]
        if not codes:                      # tables 0, 4, 14: no codes, every pair is (0, 0) using 0 bits
            base[t], rootw[t] = 0xFFFF, 0
            raw_base[t], raw_count[t] = 0, 0
            continue
        if src not in built:
            w0, entries = build_lut(codes)
            built[src] = (len(lut), w0, len(raw), len(codes))
            lut += entries
            raw += codes
        base[t], rootw[t], raw_base[t], raw_count[t] = built[src]
    # count1 table A (ISO table "32"): single 6-bit lookup, entry = len<<4 | vwxy; table B is ~bits (4 bits)
    c1 = [None]*64
    for (_, v, l, c) in tabs[32]["codes"]:
        for k in range(1 << (6 - l)):
            c1[(c << (6 - l)) + k] = (l << 4) | v
    assert all(e is not None for e in c1)
    for (_, v, l, c) in tabs[33]["codes"]:
        assert l == 4 and v == (~c & 15)
    return dict(lut=lut, base=base, rootw=rootw, raw=raw, raw_base=raw_base, raw_count=raw_count, count1a=c1)


# --------------------------------------------------------------------------- window (needs oracle/_ref)
def window_from_reference():
    """512 integer taps Dw[n] with  pcm[32 t + j] = sum_tau Dw[32 tau + j] * V_{t-tau}[j + 32 (tau & 1)],
    V_t[i] = sum_k cos(pi (2k+1)(16+i)/64) s_t[k]   (ISO 11172-3 synthesis filterbank, D scaled by 65536)."""
    from oracle.refshim import Ref
    ref = Ref(float_output=True)
    n_gr = 2
    dw = np.zeros(512)
    # impulse in subband k at slot 0; float output = pcm/32768
    resp = {}
    for k in (0, 1, 5):
        sub = np.zeros((n_gr, 1, 576), np.float32)
        sub[0, 0, k*18 + 0] = 1.0
        resp[k] = ref.synth_sequence(sub, 1).astype(np.float64)*32768.0
    for n in range(512):
        tau, j = divmod(n, 32)
        i = j + 32*(tau & 1)
        best = None
        for k in (0, 1, 5):
            c = math.cos(math.pi*(2*k + 1)*(16 + i)/64.0)
            if abs(c) > 0.3:
                v = resp[k][32*tau + j]/c
                best = v if best is None else best
                assert abs(v - best) < 0.51, (n, v, best)
        if best is None:                      # i == 16: cos == 0 for every k, tap is unobservable (ISO D[..] * 0)
            best = 0.0
        assert abs(best - round(best)) < 0.05, (n, best)
        dw[n] = round(best)
    return [int(v) for v in dw]


# --------------------------------------------------------------------------- emit
def c_array(name, ctype, vals, per_line=16, fmt="%d"):
    s = "MP3T_QUAL %s %s[%d] = {\n" % (ctype, name, len(vals))
    for i in range(0, len(vals), per_line):
        s += "    " + ",".join(fmt % v for v in vals[i:i + per_line]) + ",\n"
    return s + "};\n"


def f32(v):
    s = "%.9g" % float(np.float32(v))
    if "." not in s and "e" not in s:
        s += ".0"
    return s + "f"


def generate():
    long_t, short_t, mixed_t = sfb_tables()
    p43 = pow43_table()
    huff = huffman_tables(os.path.join(REF, "huffopt", "HUFFCODE"))
    dwin = window_from_reference()
    assert dwin[0] == 0 and dwin[64] == 213 and dwin[256] == 75038, (dwin[0], dwin[64], dwin[256])

    out = []
    out.append("/* GENERATED by tools/gen_tables.py -- do not edit.  ISO/IEC 11172-3 & 13818-3 Layer III constants. */\n"
               "#ifndef MP3_TABLES_H\n#define MP3_TABLES_H\n#include <stdint.h>\n#ifndef MP3T_QUAL\n#define MP3T_QUAL static const\n#endif\n\n")
    out.append("/* scalefactor band widths, 0-terminated; row = sample-rate class (see tools/gen_tables.py) */\n")
    out.append(c_array("MP3T_SFB_LONG", "uint8_t", [w for r in long_t for w in r], 23).replace("[184]", "[8*23]"))
    out.append(c_array("MP3T_SFB_SHORT", "uint8_t", [w for r in short_t for w in r], 40).replace("[320]", "[8*40]"))
    out.append(c_array("MP3T_SFB_MIXED", "uint8_t", [w for r in mixed_t for w in r], 40).replace("[320]", "[8*40]"))
    out.append("\n/* |q|^(4/3), q = 0..128 at [16..144]; [0..15] = -(q^(4/3)), q = 0..15 (sign-folded lookups) */\n")
    p43s = [-(float(v)) for v in p43[:16]] + [float(v) for v in p43[:129]]
    out.append("MP3T_QUAL float MP3T_POW43[145] = {\n")
    for i in range(0, 145, 8):
        out.append("    " + ",".join(f32(v) for v in p43s[i:i + 8]) + ",\n")
    out.append("};\n")
    out.append("\n/* Huffman big-value tables: multi-level LUT, %d-bit root, <=%d-bit sub-tables.\n"
               "   leaf = len<<8 | x<<4 | y ; link = 0x8000 | width<<11 | offset-from-table-base */\n" % (ROOT_BITS, SUB_BITS))
    out.append("#define MP3T_HUFF_LUT_SIZE %d\n" % len(huff["lut"]))
    out.append(c_array("MP3T_HUFF_LUT", "uint16_t", huff["lut"], 16, "0x%04x"))
    out.append(c_array("MP3T_HUFF_BASE", "uint16_t", huff["base"], 16, "%d"))
    out.append(c_array("MP3T_HUFF_ROOTW", "uint8_t", huff["rootw"], 32, "%d"))
    out.append(c_array("MP3T_HUFF_LINBITS", "uint8_t", LINBITS, 32, "%d"))
    out.append("\n/* ISO code words (for the oracle's bit-serial decoder): {len, code, x, y} */\n")
    out.append("typedef struct { uint8_t len; uint32_t code; uint8_t x, y; } mp3t_huffcode_t;\n")
    out.append("MP3T_QUAL mp3t_huffcode_t MP3T_HUFF_CODES[%d] = {\n" % len(huff["raw"]))
    for i in range(0, len(huff["raw"]), 6):
        out.append("    " + ",".join("{%d,0x%x,%d,%d}" % (l, c, x, y) for (x, y, l, c) in huff["raw"][i:i + 6]) + ",\n")
    out.append("};\n")
    out.append(c_array("MP3T_HUFF_CODES_BASE", "uint16_t", huff["raw_base"], 16, "%d"))
    out.append(c_array("MP3T_HUFF_CODES_COUNT", "uint16_t", huff["raw_count"], 16, "%d"))
    out.append("\n/* count1 table A: 6-bit lookup, len<<4 | vwxy ; table B: vwxy = ~bits & 15, len 4 */\n")
    out.append(c_array("MP3T_COUNT1A", "uint8_t", huff["count1a"], 16, "0x%02x"))
    out.append("\n/* synthesis window D[n]*65536 (integers), flat ISO order n = 0..511 */\n")
    out.append(c_array("MP3T_DWIN", "int32_t", dwin, 16, "%d"))
    out.append("\n#endif /* MP3_TABLES_H */\n")
    text = "".join(out)
    for rel in ("minimp3_b200/csrc/mp3_tables.h", "oracle/mp3_tables.h"):
        with open(os.path.join(ROOT, rel), "w") as f:
            f.write(text)
    print("wrote tables: huff lut %d entries, raw codes %d" % (len(huff["lut"]), len(huff["raw"])))
    return dict(long=long_t, short=short_t, mixed=mixed_t, pow43=p43, huff=huff, dwin=dwin)


def verify_against_reference(t):
    from oracle.refshim import Ref
    ref = Ref()
    # header words selecting each sample-rate class (layer III, no CRC): byte1 = version/layer, byte2 = sr bits
    hdrs = {0: 0xFFE30000 | (0x10 << 8) | 0xC0, 1: 0xFFE30000 | (0x18 << 8) | 0xC0,
            2: 0xFFF30000 | (0x10 << 8) | 0xC0, 3: 0xFFF30000 | (0x14 << 8) | 0xC0, 4: 0xFFF30000 | (0x18 << 8) | 0xC0,
            5: 0xFFFB0000 | (0x10 << 8) | 0xC0, 6: 0xFFFB0000 | (0x14 << 8) | 0xC0, 7: 0xFFFB0000 | (0x18 << 8) | 0xC0}
    for row, h in hdrs.items():
        for kind, tab in ((0, t["long"]), (1, t["short"]), (2, t["mixed"])):
            buf = np.zeros(64, np.uint8)
            h32 = h if h < (1 << 31) else h - (1 << 32)
            n = ref.lib.ref_const_sfb_table(h32, kind, buf.ctypes.data, 64)
            got = list(buf[:n])
            want = list(tab[row])
            want = want[:want.index(0) + 1]
            assert got == want, (row, kind, got, want)
    for i in range(129):
        assert np.float32(ref.lib.ref_const_pow43(i)) == t["pow43"][i], i
    print("tables verified against compiled reference")


if __name__ == "__main__":
    tabs = generate()
    verify_against_reference(tabs)
