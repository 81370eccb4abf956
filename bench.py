#!/usr/bin/env python3
"""bench.py -- decoded PCM samples/s of the batched Layer III hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--streams S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): S = 1024 streams per GPU, each the 44.1 kHz stereo Layer III
conformance stream l3-sin1k0db.bit (tests/golden/vectors, tiled verbatim -> "synthetic").
A step = one pass of the hot path (entropy + stereo + transform kernels) over the whole batch.

  value  device-resident: bitstreams + descriptor tables already in HBM, CUDA events on the launching
         stream around exactly K steps, max over ranks.
  e2e    the public C-ABI call mp3dec_decode_batch_cuda() with HOST buffers every step: host parse,
         pinned staging, H2D, kernels, D2H of all PCM inside the timed region.
  roofline  dominant kernel timed alone (CUDA events), algorithmic bytes / duration vs measured HBM peak.
  cpu_baseline  the UNMODIFIED reference (oracle/_ref) on all host cores, same batch, rank 0, N=1.

--impl reference times that CPU reference alone with the same metric/config.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
VECTOR_DIR = os.path.join(ROOT, "tests", "golden", "vectors")
DEFAULT_VECTOR = "l3-sin1k0db.bit"     # BASELINE.json configs[1]: 44.1 kHz stereo Layer III
METRIC = "decoded PCM samples/sec (whole box), 44.1k stereo L3 batch"
UNIT = "samples/s"
# algorithmic bytes per output sample: transform kernel 4 B spectra in + 2 B PCM out;
# entropy kernel 0.18 B bitstream in + 4 B spectra out
BYTES_PER_SAMPLE = {"transform": 6.0, "entropy": 133120.0/725760.0 + 4.0}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ts, line in self.lines:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:   # region shorter than one sample: fall back to everything we saw
            for ts, line in self.lines:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[0]))
                    mx.append(float(f[1]))
                except Exception:
                    pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, rank, world):
    """CPU arm: the unmodified reference on all host cores (oracle/_ref).  Rank 0 only."""
    if rank != 0:
        return
    from oracle import refshim
    refshim.build()
    if not refshim.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built and /root/reference absent"}))
        return
    ref = refshim.Ref()
    data = open(os.path.join(VECTOR_DIR, args.vector), "rb").read()
    vname = os.path.splitext(args.vector)[0]
    vdesc = "44.1 kHz stereo L3" if args.vector == DEFAULT_VECTOR else "non-default stream, exploration only"
    cores = os.cpu_count() or 1
    # bounded sample: enough streams for ~1 s of wall time per step, never more than the GPU batch
    n_streams = max(cores, min(args.streams, cores*32))
    streams = [data]*n_streams
    for _ in range(min(args.warmup, 2)):
        ref.bench_decode(streams[:cores*2], cores)
    steps = min(args.steps, 10)
    total, secs = 0, 0.0
    for _ in range(steps):
        s, t = ref.bench_decode(streams, cores)
        total += s
        secs += t
    value = total/secs
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 2),
            "ms_per_step": 1e3*secs/steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic: %s (Layer III conformance stream) tiled" % args.vector,
            "impl": "reference",
            "config": {"workload": "%d-stream batch tiled from %s (%s), CPU sample of the %d-stream GPU batch"
                       % (n_streams, vname, vdesc, args.streams), "streams": n_streams},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                             "sample": "%d streams x %d steps, minimp3 SSE2 path (gcc -O2), one mp3dec_t per stream, pthreads"
                             % (n_streams, steps)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=1024, help="streams per GPU")
    ap.add_argument("--vector", default=DEFAULT_VECTOR, help="stream in tests/golden/vectors the batch is tiled from (other than the default: exploration only)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import minimp3_b200 as m

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the decode path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    data = open(os.path.join(VECTOR_DIR, args.vector), "rb").read()
    vname = os.path.splitext(args.vector)[0]
    vdesc = "44.1 kHz stereo L3" if args.vector == DEFAULT_VECTOR else "non-default stream, exploration only"
    streams = [data]*args.streams          # independent streams; shards need no collective
    stream = torch.cuda.Stream()
    plan = m.BatchPlan(streams, device=local_rank, raw_frames=False, output=m.OUT_PINNED, cuda_stream=stream.cuda_stream)
    st = plan.stats
    samples_per_step = int(st["samples"])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ------------------------------------------------------------------
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            plan.run()
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        time.sleep(0.25)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t_wall0 = time.time()
        ev0.record(stream)
        for _ in range(args.steps):
            plan.run()
        ev1.record(stream)
        barrier()
        t_wall1 = time.time()
        ms = ev0.elapsed_time(ev1)
        clocks = sampler.stop(t_wall0, t_wall1)

        # ---- per-kernel durations (each stage alone, same data) -------------------------------------------
        kern = {}
        for name, stage in (("entropy", 0), ("stereo", 1), ("transform", 2)):
            if stage == 1 and not st["n_istereo_granules"]:
                continue
            plan.run_stage(stage)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            a.record(stream)
            for _ in range(reps):
                plan.run_stage(stage)
            b.record(stream)
            torch.cuda.synchronize()
            kern[name] = a.elapsed_time(b)/reps

    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    total_samples = samples_per_step*world
    value = total_samples*args.steps/(ms_max*1e-3)

    # ---- end to end through the public C-ABI call (host buffers in, host PCM out) ----------------------
    h2d = int(plan.stats["h2d_bytes"])
    parse_ms, h2d_ms = plan.stats["host_parse_ms"], plan.stats["h2d_ms"]
    plan.close()
    e2e_ms = []
    host_threads = max(2, (os.cpu_count() or 2)//max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world))))
    batch = m._Batch(streams, device=local_rank, raw_frames=False, output=m.OUT_PINNED, host_threads=host_threads)
    n_out = 0
    for i in range(args.e2e_steps + 2):
        barrier()
        t0 = time.perf_counter()
        n_out = batch.call()          # host parse + pinned staging + H2D + kernels + D2H of all PCM, blocking
        t1 = time.perf_counter()
        if i >= 2:
            e2e_ms.append((t1 - t0)*1e3)
    rets = list(batch.rets)[:batch.n]
    assert all(r == 0 for r in rets) and n_out == samples_per_step
    te = torch.tensor([statistics.median(e2e_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = n_out*world/(float(te.item())*1e-3)

    # ---- same call, PCM left in HBM (SURVEY 8d timing scope ii: host parse + H2D + kernels, no D2H) ------
    dev_ms = []
    batch_dev = m._Batch(streams, device=local_rank, raw_frames=False, output=m.OUT_DEVICE, host_threads=host_threads)
    for i in range(args.e2e_steps + 2):
        barrier()
        t0 = time.perf_counter()
        n_dev = batch_dev.call()
        t1 = time.perf_counter()
        if i >= 2:
            dev_ms.append((t1 - t0)*1e3)
    assert n_dev == samples_per_step
    td = torch.tensor([statistics.median(dev_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(td, op=dist.ReduceOp.MAX)
    dev_value = n_dev*world/(float(td.item())*1e-3)

    if rank == 0:
        peak, peak_kind = measured_peaks()
        dom = max(("entropy", "transform"), key=lambda k: kern.get(k, 0.0))
        achieved = BYTES_PER_SAMPLE[dom]*samples_per_step/(kern[dom]*1e-3)/1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(dom)
            except Exception:
                traffic = None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max/args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic: %s (Layer III conformance stream) tiled" % args.vector,
                "config": {"workload": "%d-stream batch tiled from %s (%s) per GPU" % (args.streams, vname, vdesc),
                           "streams_per_gpu": args.streams, "frames_per_gpu": int(st["n_frames"]),
                           "granule_channels_per_gpu": int(st["n_granule_channels"]), "samples_per_step_per_gpu": samples_per_step,
                           "parallelism": "streams sharded over %d GPU(s), no collective" % world,
                           "l2": "working set %.0f MB per step (bitstream + spectra + PCM) > 126 MB L2, no flush needed"
                                 % ((st["input_bytes"] + 4*576*st["n_granule_channels"] + 2*samples_per_step)/1e6)},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 2*n_out,
                        "ms_per_step": float(te.item()), "host_parse_ms": parse_ms, "h2d_ms": h2d_ms,
                        "host_threads_per_rank": host_threads,
                        "api": "mp3dec_decode_batch_cuda(host buffers) -> pinned host PCM, chunk-pipelined inside the call",
                        "pcm_left_on_device": {"value": dev_value, "unit": UNIT, "ms_per_step": float(td.item()),
                                               "note": "same call with MP3D_BATCH_OUT_DEVICE: host parse + H2D + kernels, no D2H"}},
                "gpu_launches": int(st["kernels_per_run"])*args.steps,
                "kernels_ms": kern,
                "roofline": {"bound": "hbm", "kernel": "k_l3_" + dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved/peak, "traffic": traffic, "peak_kind": peak_kind,
                             "bytes_per_sample": BYTES_PER_SAMPLE[dom]},
                "clocks": clocks}
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle import refshim
                refshim.build()
                ref = refshim.Ref()
                cores = os.cpu_count() or 1
                n_cpu = max(cores, min(args.streams, cores*32))
                ref.bench_decode(streams[:cores], cores)
                s_cpu, t_cpu = ref.bench_decode(streams[:n_cpu], cores)
                line["cpu_baseline"] = {"value": s_cpu/t_cpu, "unit": UNIT, "cores": cores, "kind": "reference",
                                        "sample": "%d of the %d streams, 1 pass, unmodified minimp3 (SSE2, gcc -O2), pthreads"
                                                  % (n_cpu, args.streams)}
            except Exception as e:  # the checker being absent must not hide the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
