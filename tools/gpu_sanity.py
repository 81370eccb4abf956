#!/usr/bin/env python3
"""Small all-paths run for compute-sanitizer: batch call over Layer I/II/III streams (MS, intensity stereo, short and
mixed blocks, LSF, float output) plus the single-frame API, each checked against the compiled reference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import minimp3_b200 as m  # noqa: E402
from oracle.refshim import Ref  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "vectors")
names = ["l3-si_block", "l3-he_mode", "l3-si_huff", "M2L3_bitrate_16_all", "l1-fl4", "l2-fl13", "l3-hecommon"]
datas = [open(os.path.join(GOLD, n + ".bit"), "rb").read() for n in names]
ref = Ref()
rets, outs = m.mp3dec_decode_batch_cuda(datas, raw_frames=True)
for n, d, r, o in zip(names, datas, rets, outs):
    want, _ = ref.decode_frames(d)
    diff = int(np.abs(o["pcm"].astype(np.int32) - want.astype(np.int32)).max())
    assert r == 0 and len(want) == len(o["pcm"]) and diff <= 1, (n, r, diff)
rets, outs = m.mp3dec_decode_batch_cuda(datas[:3], raw_frames=True, float_output=True)
pcm, frames = m.decode_frames(datas[0])
want, _ = ref.decode_frames(datas[0])
assert np.abs(pcm.astype(np.int32) - want.astype(np.int32)).max() <= 1
pcm, frames = m.decode_frames(datas[5])
print("sanity ok")
