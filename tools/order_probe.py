#!/usr/bin/env python3
"""Entropy-kernel time of a mixed batch (every Layer III vector of tests/golden/vectors, repeated and interleaved so
that the 32 streams of a window are mostly different) under the three lane-order policies of k_order.cu:
MP3B_DEBUG_ORDER_ALIKE=0 per-stream order, =1 every window merged, unset: merged only where streams are alike."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    import minimp3_b200 as m
    vdir = os.path.join(ROOT, "tests/golden/vectors")
    names = sorted(n for n in os.listdir(vdir) if n.startswith(("l3-", "M2L3")))
    data = [open(os.path.join(vdir, n), "rb").read() for n in names]
    streams = (data*24)[:len(data)*24]
    plan = m.BatchPlan(streams, device=0, raw_frames=False, output=m.OUT_PINNED)
    for _ in range(3):
        plan.run_stage(0)
    torch.cuda.synchronize()
    plan.sync()
    import time
    t0 = time.perf_counter()
    for _ in range(20):
        plan.run_stage(0)
    plan.sync()
    print("%s: %d streams, %d granule-channels, entropy %.3f ms" % (os.environ.get("MP3B_DEBUG_ORDER_ALIKE", "auto"), len(streams),
          int(plan.stats["n_granule_channels"]), (time.perf_counter() - t0)*1e3/20))
    plan.close()
else:
    for mode in ("0", "1", None):
        env = dict(os.environ)
        env.pop("MP3B_DEBUG_ORDER_ALIKE", None)
        if mode is not None:
            env["MP3B_DEBUG_ORDER_ALIKE"] = mode
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, check=True)
