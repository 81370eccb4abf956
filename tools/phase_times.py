#!/usr/bin/env python3
"""Per-phase durations of k_l3_transform from in-kernel clock64 stamps (debug taps), config-2 style batch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MP3B_DEBUG_CLK_ONLY", "1")
import minimp3_b200 as m  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
data = open(os.path.join(ROOT, "tests", "golden", "vectors", "l3-sin1k0db.bit"), "rb").read()
p = m.BatchPlan([data]*n)
p.enable_taps()
p.run()
p.sync()
p.run()
p.sync()
clk = p.read_tap(7)
p.close()
names = ["p0 desc+zero+tma issue", "p1 tma wait", "p1 ms", "p1 barrier", "phase3a imdct core", "phase3b overlap",
         "phase4 dct", "phase5 fir + special round", "end"]
valid = clk[:, 0] > 0
c = clk[valid]
nst = int((c[0] > 0).sum())
d = np.diff(c[:, :nst], axis=1).astype(np.float64)
tot = (c[:, nst - 1] - c[:, 0]).astype(np.float64)
print("tiles", len(c), "stamps", nst, "mean tile cycles", tot.mean(), "=", tot.mean()/1.965e3, "us")
for i in range(nst - 1):
    print("%-22s mean %8.0f cyc  (%5.1f%%)  p90 %8.0f" % (names[i] if i < len(names) else i, d[:, i].mean(), 100*d[:, i].mean()/tot.mean(),
                                                    np.percentile(d[:, i], 90)))
