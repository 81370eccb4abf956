#!/usr/bin/env python3
"""Developer aid: decode the golden vectors on the GPU and print parity against the compiled reference
(oracle/_ref), stage by stage.  Needs a B200 (run under gpurun)."""
import glob
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import minimp3_b200 as m  # noqa: E402
from oracle.refshim import Ref  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "vectors")


def main():
    names = sys.argv[1:] or ["l3-sin1k0db", "l3-compl", "l3-si_huff", "l3-si_block", "l3-he_mode", "M2L3_bitrate_16_all",
                             "l3-he_free", "l3-test45", "M2L3_noise", "l3-hecommon"]
    ref = Ref()
    print(m.load().mp3dec_b200_version())
    for name in names:
        data = open(os.path.join(GOLD, name + ".bit"), "rb").read()
        t = ref.tap_stream(data)
        t0 = time.time()
        p = m.BatchPlan([data], raw_frames=True)
        p.enable_taps()
        p.run()
        rets, outs = p.fetch()
        dt = time.time() - t0
        pcm = outs[0]["pcm"]
        desc = p.read_tap(4)
        q = p.read_tap(0)
        scf = p.read_tap(1)
        xr = p.read_tap(2)
        sub = p.read_tap(3)
        nzl = p.read_tap(6)
        p.close()
        # real granule-channels = those with non-dummy descriptors, in order: match by count
        flags1 = desc[:, 29]
        p23 = desc[:, 8].astype(int) | (desc[:, 9].astype(int) << 8)
        # dummies have all-zero body and flags1 == 1
        body = desc[:, :28].astype(int).sum(axis=1)
        real = np.where(~((body == 0) & (flags1 == 1)))[0]
        msg = [name, "ret", rets[0], "samples", len(pcm), len(t["pcm"]), "gcs", len(real), len(t["rec"]), "%.1f ms" % (dt*1e3)]
        if len(real) == len(t["rec"]):
            msg += ["q", bool(np.array_equal(q[real], t["q"]))]
            nsfb = np.array([r.n_long_sfb + r.n_short_sfb for r in t["rec"]])
            ok = all(np.array_equal(scf[real[i], :nsfb[i]].view(np.uint32), t["scf"][i, :nsfb[i]].view(np.uint32))
                     for i in range(len(real)))
            msg += ["scf", ok]
            xrz = xr[real].copy()
            for i, g in enumerate(real):
                xrz[i, nzl[g]:] = 0
            ms_plain = (desc[real, 29] & 0x10) != 0
            want = np.where(ms_plain[:, None], t["xr"], t["xr_stereo"])
            msg += ["xr", bool(np.array_equal(xrz.view(np.uint32), want.view(np.uint32)))]
            scale = max(1.0, float(np.abs(t["sub"]).max()))
            msg += ["sub_relerr", float(np.abs(sub[real] - t["sub"]).max()/scale)]
        if len(pcm) == len(t["pcm"]):
            d = pcm.astype(int) - t["pcm"].astype(int)
            msg += ["pcm_maxdiff", int(np.abs(d).max()) if len(d) else 0, "frac", float((d != 0).mean()) if len(d) else 0.0]
        print(*msg, flush=True)


if __name__ == "__main__":
    main()
