#!/usr/bin/env python3
"""Times the one-call entry point on a config (host buffers in): where the call's time goes, per output mode, with the
GPU frame scan off / on, and with the streams in a caller-pinned slab.   python tools/e2e_probe.py [config] [calls]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import minimp3_b200 as m  # noqa: E402
from minimp3_b200 import corpus  # noqa: E402


def pinned_slab(streams):
    """all streams back to back (16-byte aligned) in one page-locked buffer -> (views, (address, bytes), keep-alive)"""
    import torch
    offs, total = [], 0
    for s in streams:
        offs.append(total)
        total += (len(s) + 15) & ~15
    t = torch.empty(max(total, 16), dtype=torch.uint8, pin_memory=True)
    a = t.numpy()
    views = []
    for s, o in zip(streams, offs):
        a[o:o + len(s)] = np.frombuffer(s, dtype=np.uint8)
        views.append(a[o:o + len(s)])
    return views, (a.ctypes.data, total), t


if __name__ == "__main__":
    cfg = sys.argv[1] if len(sys.argv) > 1 else "2"
    calls = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    streams, _ = corpus.build_streams(cfg)
    views, slab, keep = pinned_slab(streams)
    for label, scan, src, extra in (("host walk", -1, streams, {}), ("gpu scan", 1, streams, {}),
                                    ("gpu scan, pinned inputs", 1, views, {"pinned_inputs": slab})):
        for name, out in (("device", m.OUT_DEVICE), ("pinned", m.OUT_PINNED)):
            b = m._Batch(src, allow_mono_stereo=True, output=out, gpu_frame_scan=scan, **extra)
            ts = []
            for i in range(calls):
                t0 = time.perf_counter()
                n = b.call()
                ts.append((time.perf_counter() - t0)*1e3)
            cs = b.call_stats()
            print("cfg %s %-24s out %-6s: %.2f ms (best %.2f) = %.1f Gsamples/s; scanned %d, host %.2f ms, issue %.2f, finish %.2f; calls %s"
                  % (cfg, label, name, sorted(ts)[len(ts)//2], min(ts), n/min(ts)/1e6, cs["scanned_streams"], cs["host_parse_ms"],
                     cs["issue_ms"], cs["finish_ms"], " ".join("%.0f" % t for t in ts)), flush=True)
