#!/usr/bin/env python3
"""Build tests/golden/ from the reference checkout (run in the build container only).

* copies the conformance / regression bitstreams (data fixtures, not source) the
  hot path is pinned by into tests/golden/vectors/
* stores the ISO / regression PCM that ships next to them, xz-compressed, under
  tests/golden/pcm/
* records in tests/golden/manifest.json, per vector, what the UNMODIFIED reference
  (oracle/_ref) produced here: frame count, sample count, sha256 of its int16 PCM,
  max |diff| and PSNR against the sibling .pcm  (the quantities minimp3_test.c:391-428
  prints) and the result of mp3dec_load_buf (return code, samples, sha256).

/root/reference does not exist on the GPU box; everything tests need at run time
lands in tests/golden/.
"""
import hashlib
import json
import lzma
import math
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.refshim import Ref, build  # noqa: E402

REF = os.environ.get("MINIMP3_REFERENCE", "/root/reference")
VEC = os.path.join(REF, "vectors")
OUT = os.path.join(ROOT, "tests", "golden")

L3 = ["l3-compl", "l3-he_32khz", "l3-he_44khz", "l3-he_48khz", "l3-he_free", "l3-he_mode", "l3-hecommon",
      "l3-si", "l3-si_block", "l3-si_huff", "l3-sin1k0db", "l3-test45", "l3-test46",
      "M2L3_bitrate_16_all", "M2L3_bitrate_22_all", "M2L3_bitrate_24_all", "M2L3_compl24", "M2L3_noise"]
NONSTD = ["l3-nonstandard-apetag", "l3-nonstandard-big-iscf", "l3-nonstandard-compl-sideinfo-bigvalues",
          "l3-nonstandard-compl-sideinfo-blocktype", "l3-nonstandard-compl-sideinfo-size",
          "l3-nonstandard-he_44_48khz", "l3-nonstandard-id3v1-apetag", "l3-nonstandard-id3v1",
          "l3-nonstandard-id3v2-only", "l3-nonstandard-id3v2", "l3-nonstandard-sideinfo-size",
          "l3-nonstandard-sin1k0db_lame_vbrtag", "l3-nonstandard-small", "l3-nonstandard-vbrtag-corrupted",
          "l3-nonstandard-vbrtag-empty", "l3-nonstandard-vbrtag-noframes", "l3-nonstandard-vbrtag-only",
          "l3-nonstandard-vbrtag-oob-read"]
L12 = ["l1-fl1", "l1-fl2", "l1-fl3", "l1-fl4", "l1-fl5", "l1-fl6", "l1-fl7", "l1-fl8",
       "l2-fl10", "l2-fl11", "l2-fl12", "l2-fl13", "l2-fl14", "l2-fl15", "l2-fl16", "l2-test32",
       "l2-nonstandard-fl1_fl2_ff", "l2-nonstandard-free_format", "l2-nonstandard-test32-size",
       "ILL2_center2", "ILL2_dual", "ILL2_dynx22", "ILL2_dynx31", "ILL2_dynx32", "ILL2_ext_switching", "ILL2_layer1",
       "ILL2_layer3", "ILL2_mono", "ILL2_multilingual", "ILL2_overalloc1", "ILL2_overalloc2", "ILL2_prediction",
       "ILL2_samples", "ILL2_scf63", "ILL2_tca21", "ILL2_tca30", "ILL2_tca30_PC", "ILL2_tca31_PC", "ILL2_tca31_mtx0",
       "ILL2_tca31_mtx2", "ILL2_tca32_PC", "ILL2_wrongcrc", "ILL4_ext_id1", "ILL4_sync", "ILL4_wrong_length1",
       "ILL4_wrong_length2", "ILL4_wrongcrc"]
PERF = ["performance/MIPSTest.mp3"]


def sha(b):
    return hashlib.sha256(b).hexdigest()


def psnr(a, b):
    n = min(len(a), len(b))
    if n == 0:
        return 0, 99.0
    d = np.abs(a[:n].astype(np.int64) - b[:n].astype(np.int64))
    mse = float((d.astype(np.float64)**2).sum())/max(1, len(a))
    return int(d.max()), (99.0 if mse == 0 else 10.0*math.log10((32767.0*32767.0)/mse))


def main():
    build()
    r16 = Ref(False)
    os.makedirs(os.path.join(OUT, "vectors"), exist_ok=True)
    os.makedirs(os.path.join(OUT, "pcm"), exist_ok=True)
    manifest = {}
    for group, names in (("l3", L3), ("nonstandard", NONSTD), ("l12", L12), ("perf", PERF)):
        for name in names:
            src = os.path.join(VEC, name if group == "perf" else name + ".bit")
            base = os.path.basename(name)
            dst = os.path.join(OUT, "vectors", base if group == "perf" else base + ".bit")
            shutil.copyfile(src, dst)
            data = open(src, "rb").read()
            pcm, frames = r16.decode_frames(data)
            ret, lpcm, linfo = r16.load_buf(data)
            e = dict(group=group, file=os.path.basename(dst), bytes=len(data), sha256=sha(data),
                     frames=len(frames), decoded_frames=sum(1 for f in frames if f.samples),
                     samples=int(len(pcm)), pcm_sha256=sha(pcm.tobytes()),
                     load_ret=ret, load_samples=int(len(lpcm)), load_pcm_sha256=sha(lpcm.tobytes()), load_info=linfo,
                     hz=frames[0].hz if frames else 0, layers=sorted({f.layer for f in frames if f.samples}),
                     channels=sorted({f.channels for f in frames if f.samples}))
            pcm_src = os.path.join(VEC, name + ".pcm")
            if group != "perf" and os.path.exists(pcm_src):
                ref_pcm = np.fromfile(pcm_src, dtype="<i2")
                with lzma.open(os.path.join(OUT, "pcm", base + ".pcm.xz"), "wb", preset=9) as f:
                    f.write(ref_pcm.tobytes())
                md, ps = psnr(lpcm, ref_pcm)
                e.update(ref_pcm_samples=int(len(ref_pcm)), ref_max_diff=md, ref_psnr=round(ps, 6))
            manifest[base] = e
            print(base, e["frames"], e["samples"], e.get("ref_max_diff"), e.get("ref_psnr"), ret)
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
