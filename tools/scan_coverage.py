#!/usr/bin/env python3
"""Which conformance vectors the GPU frame scan (k_scan.cu) takes and which it hands back to the host walker.
    python tools/scan_coverage.py        (needs a GPU)"""
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import minimp3_b200 as m  # noqa: E402

if __name__ == "__main__":
    taken, back = [], []
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "vectors", "*"))):
        name = os.path.basename(path)
        if not name.endswith((".bit", ".mp3")):
            continue
        data = open(path, "rb").read()
        b = m._Batch([data]*64, allow_mono_stereo=True, output=m.OUT_DEVICE, gpu_frame_scan=1)
        b.call()
        n = b.call_stats()["scanned_streams"]
        (taken if n == 64 else back).append("%s (%d)" % (name, n) if n not in (0, 64) else name)
        del b
    print("scanned on the GPU (%d): %s" % (len(taken), ", ".join(taken)))
    print("handed back to the host walker (%d): %s" % (len(back), ", ".join(back)))
