#!/usr/bin/env python3
"""What the box delivers for device-to-host copies, alone and with N GPUs copying at once: the floor of `e2e`
(PCM read-back) that no decoder can beat.  One thread per GPU, 1.5 GB device buffer -> its own pinned host buffer,
cudaMemcpyAsync on a private stream, CUDA-event timed, all GPUs started together.
    python tools/d2h_ceiling.py            # uses every visible GPU: N = 1, 2, 4, 8"""
import threading
import time

import torch

n_all = torch.cuda.device_count()
size = 1_486_356_480       # bytes of PCM of one config-2 step
bufs = []
for d in range(n_all):
    with torch.cuda.device(d):
        dev = torch.empty(size, dtype=torch.uint8, device="cuda")
        host = torch.empty(size, dtype=torch.uint8, pin_memory=True)
        bufs.append((dev, host, torch.cuda.Stream(device=d)))
for n in [k for k in (1, 2, 4, 8) if k <= n_all]:
    barrier = threading.Barrier(n)
    times = [0.0]*n

    def work(d):
        dev, host, st = bufs[d]
        torch.cuda.set_device(d)
        with torch.cuda.stream(st):
            host.copy_(dev, non_blocking=True)
            st.synchronize()
            barrier.wait()
            t0 = time.perf_counter()
            for _ in range(4):
                host.copy_(dev, non_blocking=True)
            st.synchronize()
            times[d] = time.perf_counter() - t0

    th = [threading.Thread(target=work, args=(d,)) for d in range(n)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    worst = max(times)
    print("N=%d: %.1f GB/s in aggregate (%.1f GB/s per GPU), %.1f ms for one 1.49 GB read-back per GPU"
          % (n, n*4*size/worst/1e9, 4*size/worst/1e9, worst/4*1e3), flush=True)
