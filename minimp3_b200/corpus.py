"""Workloads of BASELINE.json `configs` as lists of stream buffers.

Used by bench.py (--config) and by the full-size parity tests.  Streams are tiled verbatim from the bundled
conformance vectors (tests/golden/vectors); every stream gets its OWN host buffer (no aliasing of one bytes object),
so host walks and staging copies see a batch-sized working set.
"""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VECTOR_DIR = os.path.join(ROOT, "tests", "golden", "vectors")

L1 = ["l1-fl%d.bit" % i for i in range(1, 9)]
L2 = ["l2-fl%d.bit" % i for i in range(10, 17)]
L3 = ["l3-compl.bit", "l3-he_32khz.bit", "l3-he_44khz.bit", "l3-he_48khz.bit", "l3-he_free.bit", "l3-he_mode.bit",
      "l3-hecommon.bit", "l3-si.bit", "l3-si_block.bit", "l3-si_huff.bit", "l3-sin1k0db.bit", "l3-test45.bit",
      "l3-test46.bit"]

# name -> (vectors cycled stream by stream, streams per GPU, one-line description)
CONFIGS = {
    "2": (["l3-sin1k0db.bit"], 1024,
          "1024-stream batch tiled from l3-sin1k0db (44.1 kHz stereo L3)"),
    "2b": (["l3-sin1k0db.bit", "l3-he_44khz.bit"], 1024,
           "1024-stream batch, even streams l3-sin1k0db (stereo), odd streams l3-he_44khz (mono VBR), 44.1 kHz L3"),
    "3": (["l3-si_block.bit", "l3-he_mode.bit", "l3-si_huff.bit"], 3072,
          "mixed-block batch: l3-si_block / l3-he_mode / l3-si_huff cycled, 1024 streams each "
          "(short + mixed blocks, MS / intensity stereo, all Huffman tables)"),
    "4": (["M2L3_bitrate_16_all.bit", "M2L3_bitrate_22_all.bit", "M2L3_bitrate_24_all.bit", "M2L3_compl24.bit",
           "l3-he_free.bit"], 5120,
          "MPEG-2/2.5 LSF batch: M2L3_bitrate_16/22/24_all, M2L3_compl24 and free-format l3-he_free cycled, "
          "1024 streams each"),
    "5": (L1 + L2 + L3, 8192,
          "mixed Layer I/II/III corpus (l1-fl1..8, l2-fl10..16, 13 l3-* conformance streams cycled): 65536 streams "
          "over 8 GPUs = 8192 streams per GPU"),
    "mips": (["MIPSTest.mp3"], 1024,
             "1024-stream batch tiled from MIPSTest.mp3 (48 kHz stereo 320 kbit/s, dense spectra; stress case)"),
}
DEFAULT_CONFIG = "2"


def load_vector_file(name):
    with open(os.path.join(VECTOR_DIR, name), "rb") as f:
        return f.read()


def stream_names(config, n_streams=None, first=0):
    """vector name of streams first .. first + n - 1 of the (conceptually endless) cycle"""
    names, n_default, _ = CONFIGS[config]
    n = n_default if n_streams is None else int(n_streams)
    return [names[(first + i) % len(names)] for i in range(n)]


def build_streams(config, n_streams=None, first=0):
    """-> (list of bytes objects, one private copy per stream; list of vector names)"""
    names = stream_names(config, n_streams, first)
    cache = {}
    out = []
    for n in names:
        if n not in cache:
            cache[n] = load_vector_file(n)
        out.append(bytes(bytearray(cache[n])))     # a fresh buffer per stream
    return out, names


def describe(config, n_streams=None):
    names, n_default, desc = CONFIGS[config]
    n = n_default if n_streams is None else int(n_streams)
    if n != n_default:
        desc += " [run with %d streams per GPU]" % n
    return desc
