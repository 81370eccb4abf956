#!/usr/bin/env python3
"""bench.py -- decoded PCM samples/s of the batched MPEG audio hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|2b|3|4|5|mips]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload: BASELINE.json configs (minimp3_b200/corpus.py); the default and the driver's line is configs[1]:
1024 streams per GPU, each a private copy of the 44.1 kHz stereo Layer III conformance stream l3-sin1k0db.bit
(tiled verbatim -> "synthetic").  A step = one pass of the hot path over the whole batch.

  value     every GPU kernel of the job (side-info parse, lane order, entropy, stereo, tile meta, transform, Layer
            I/II) with bitstreams + frame records already in HBM; CUDA events on the launching stream around
            exactly K steps, max over ranks.                                       ("value_scope")
  scope_ii  SURVEY 8d timing scope (ii): the public call with HOST buffers, PCM left in HBM
            (host frame walk + pinned staging + H2D + kernels).
  e2e       the public C-ABI call mp3dec_decode_batch_cuda() with HOST buffers every step: host walk, pinned
            staging, H2D, kernels, D2H of all PCM inside the timed region (pinned output slab); with N > 1 ONE call
            on rank 0 drives all N GPUs (opts.devices).  e2e_malloc: the same with per-stream malloc()ed output;
            inputs sit in a caller-pinned slab (opts.pinned_inputs: direct upload, GPU frame scan); *.pageable_inputs: the
            same call on ordinary pageable buffers (host frame walk + staging copy).
  roofline  dominant kernel timed alone (CUDA events), algorithmic bytes / duration vs measured HBM peak, plus
            the FP32 figure (fast-algorithm flops / duration vs the FFMA peak at the sampled SM clock).
  cpu_baseline  the UNMODIFIED reference (oracle/_ref) on all host cores, a bounded sample of the same batch.

--impl reference times that CPU reference alone on the same streams, steps and warm-up.
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

from minimp3_b200 import corpus   # noqa: E402  (workload definitions only: no CUDA, no oracle)

METRIC = "decoded PCM samples/sec (whole box), 44.1k stereo L3 batch"
UNIT = "samples/s"
# algorithmic bytes per output sample: transform kernel 4 B spectra in + 2 B PCM out;
# entropy kernel: bitstream in (measured per config: input bytes / samples) + 4 B spectra out
TRANSFORM_BYTES_PER_SAMPLE = 6.0
# fast-algorithm flops per output sample of the transform kernel (window 32, DCT-II 9, IMDCT 12,
# antialias 3): what the FP32 pipe has to do at least
TRANSFORM_FLOPS_PER_SAMPLE = 56.0
FP32_LANES = 148*128             # FFMA lanes of one B200


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


def config_dict(args, world):
    """identical in both arms (the driver compares them)"""
    return {"workload": corpus.describe(args.config, args.streams) + ", per GPU", "config_id": args.config,
            "streams_per_gpu": args.streams, "distinct_host_buffers": True,
            "parallelism": "streams sharded over %d GPU(s), no collective" % world,
            "l2": "one step moves far more than the 126 MB L2 (bitstreams + spectra + PCM; see workload_stats), no flush needed"}


def data_string(args):
    return "synthetic: %s tiled verbatim (bundled conformance streams)" % "/".join(
        n.replace(".bit", "") for n in corpus.CONFIGS[args.config][0][:6]) + ("..." if len(corpus.CONFIGS[args.config][0]) > 6 else "")


def run_reference(args, rank, world):
    """CPU arm: the unmodified reference on all host cores (oracle/_ref), same streams, steps and warm-up as the
    GPU arm's per-GPU batch.  Rank 0 only."""
    if rank != 0:
        return
    from oracle import refshim
    refshim.build()
    if not refshim.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built and /root/reference absent"}))
        return
    ref = refshim.Ref()
    streams, _ = corpus.build_streams(args.config, args.streams)
    cores = os.cpu_count() or 1
    for _ in range(args.warmup):
        ref.bench_decode(streams, cores)
    total, secs = 0, 0.0
    per_step = []
    for _ in range(args.steps):
        s, t = ref.bench_decode(streams, cores)
        total += s
        secs += t
        per_step.append(t)
    value = total/secs
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3*secs/args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": data_string(args), "impl": "reference",
            "config": config_dict(args, world),
            "samples_per_step": total//max(1, args.steps),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                             "sample": "all %d streams of one GPU's batch x %d steps (+%d warm-up), unmodified minimp3 "
                                       "(SSE2 path, gcc -O2), mp3dec_init + mp3dec_decode_frame loop, one mp3dec_t per stream, "
                                       "%d pthreads; samples/s does not depend on the batch size"
                                       % (len(streams), args.steps, args.warmup, cores)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=corpus.DEFAULT_CONFIG, choices=sorted(corpus.CONFIGS))
    ap.add_argument("--streams", type=int, default=None, help="streams per GPU (default: the config's own count)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="kernels only (profiling runs)")
    args = ap.parse_args()
    if args.streams is None:
        args.streams = corpus.CONFIGS[args.config][1]
    if args.impl == "ours":
        args.warmup = max(args.warmup, 3)

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
    cpu_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")      # host-side barrier: idle ranks must not spin on their GPU

    # rank r decodes streams r*S .. r*S + S - 1 of the cycle: independent streams, shards need no collective
    streams, names = corpus.build_streams(args.config, args.streams, first=rank*args.streams)
    stream = torch.cuda.Stream()
    # load_buf semantics per stream (tags, VBR-tag trimming); mono/stereo transitions allowed as in the reference's own test build
    plan = m.BatchPlan(streams, device=local_rank, raw_frames=False, allow_mono_stereo=True, output=m.OUT_PINNED,
                       cuda_stream=stream.cuda_stream)
    st = plan.stats
    samples_per_step = int(st["samples"])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        plan.run_stage(4)      # GPU preparation: side-info parse + lane order
        plan.run()             # entropy, intensity stereo, tile meta, transform, Layer I/II

    # ---- device-resident throughput ------------------------------------------------------------------
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step()
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        time.sleep(0.25)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t_wall0 = time.time()
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
        barrier()
        t_wall1 = time.time()
        ms = ev0.elapsed_time(ev1)
        clocks = sampler.stop(t_wall0, t_wall1)

        # ---- per-kernel durations (each stage alone, same data) -------------------------------------------
        kern = {}
        for name, stage in (("prep", 4), ("entropy", 0), ("stereo", 1), ("transform", 2), ("layer12", 3)):
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
    tot = torch.tensor([float(samples_per_step)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_max = float(t.item())
    total_samples = int(tot.item())
    value = total_samples*args.steps/(ms_max*1e-3)

    h2d = int(plan.stats["h2d_bytes"])
    plan_parse_ms, plan_h2d_ms = plan.stats["host_parse_ms"], plan.stats["h2d_ms"]
    plan.close()

    # ---- through the public C-ABI call with host buffers -----------------------------------------------
    n_dev = world
    cores = os.cpu_count() or 2

    def timed_calls(batch, n_expect):
        """median wall time of args.e2e_steps blocking calls (2 untimed first), call stats of the last"""
        ms_list = []
        for i in range(args.e2e_steps + 2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n_out = batch.call()
            t1 = time.perf_counter()
            if i >= 2:
                ms_list.append((t1 - t0)*1e3)
            cs = batch.call_stats()
            batch.free_outputs()
            assert n_out == n_expect, (n_out, n_expect)
        assert all(r == 0 for r in list(batch.rets)[:batch.n])
        return statistics.median(ms_list), cs

    e2e = scope_ii = e2e_malloc = None
    if not args.no_e2e:
        # with N > 1 rank 0 makes ONE call that shards all N x S streams over the N GPUs (opts.devices, LPT on
        # input bytes, one shared host pool); the other ranks wait in a host-side barrier
        all_streams = streams
        if world > 1 and rank == 0:
            all_streams = []
            for r in range(world):
                all_streams += corpus.build_streams(args.config, args.streams, first=r*args.streams)[0]
        results = {}
        if world > 1:
            dist.barrier(group=cpu_group)
        if rank == 0:
            devs = list(range(n_dev))
            # primary numbers: the step's inputs live in pinned host memory (the contract's "host->device copy of that
            # step's inputs from pinned host memory"): one page-locked slab, streams back to back, announced through
            # opts.pinned_inputs -> uploaded where they are, frames found by the GPU scan (k_scan.cu), the host only
            # looks at the stream heads.  Beside them the same call on ordinary pageable buffers (host frame walk
            # into the library's own pinned staging slab, H2D from there): the plain drop-in use.
            offs, slab_bytes = [], 0
            for s_ in all_streams:
                offs.append(slab_bytes)
                slab_bytes += (len(s_) + 15) & ~15
            slab_t = torch.empty(max(slab_bytes, 16), dtype=torch.uint8, pin_memory=True)
            slab = slab_t.numpy()
            views = []
            for s_, o in zip(all_streams, offs):
                slab[o:o + len(s_)] = np.frombuffer(s_, dtype=np.uint8)
                views.append(slab[o:o + len(s_)])
            pin = (slab.ctypes.data, slab_bytes)
            for key, output, src, extra in (("e2e", m.OUT_PINNED, views, {"pinned_inputs": pin}),
                                            ("scope_ii", m.OUT_DEVICE, views, {"pinned_inputs": pin}),
                                            ("e2e_malloc", m.OUT_MALLOC, views, {"pinned_inputs": pin}),
                                            ("e2e_pageable", m.OUT_PINNED, all_streams, {}),
                                            ("scope_ii_pageable", m.OUT_DEVICE, all_streams, {})):
                batch = m._Batch(src, device=local_rank, devices=devs if n_dev > 1 else None, raw_frames=False,
                                 allow_mono_stereo=True, output=output, host_threads=0, **extra)
                ms_call, cs = timed_calls(batch, total_samples)
                results[key] = (ms_call, cs)
                del batch
        if world > 1:
            dist.barrier(group=cpu_group)
        if rank == 0:
            def pack(key, api, d2h):
                ms_call, cs = results[key]
                return {"value": total_samples/(ms_call*1e-3), "unit": UNIT,
                        "h2d_bytes_per_step": int(cs["h2d_bytes"]), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": ms_call, "host_parse_ms": cs["host_parse_ms"], "issue_ms": cs["issue_ms"],
                        "finish_ms": cs["finish_ms"], "chunks": cs["n_chunks"], "devices": cs["n_devices"],
                        "streams_frame_scanned_on_gpu": cs["scanned_streams"], "host_threads": cores, "api": api}
            e2e = pack("e2e", "ONE mp3dec_decode_batch_cuda(host buffers) call over %d GPU(s) -> pinned host PCM; streams in a "
                              "caller-pinned slab (opts.pinned_inputs): stream heads on the host, H2D, GPU frame scan + all kernels, "
                              "D2H of all PCM inside the timed call (chunk-pipelined)" % n_dev,
                       2*total_samples)
            e2e["pageable_inputs"] = pack("e2e_pageable", "same call on ordinary pageable stream buffers: host frame walk + staging copy "
                                          "into the library's pinned slab, H2D, kernels, D2H", 2*total_samples)
            e2e["d2h_gbs"] = 2*total_samples/(results["e2e"][0]*1e-3)/1e9
            e2e["note"] = ("floor = PCM bytes / device-to-host bandwidth: %.2f GB per step; tools/d2h_ceiling.py (plain cudaMemcpyAsync "
                           "into pinned memory, all GPUs at once) measured 56 / 70 / 72 / 91 GB/s in aggregate at 1 / 2 / 4 / 8 GPUs "
                           "on this pool's 8-GPU box" % (2*total_samples/1e9))
            scope_ii = pack("scope_ii", "same call with MP3D_BATCH_OUT_DEVICE (SURVEY 8d scope ii): stream heads on the host, H2D, GPU frame "
                                        "scan + kernels, PCM stays in HBM", 0)
            scope_ii["pageable_inputs"] = pack("scope_ii_pageable", "pageable stream buffers: host frame walk + staging copy + H2D + kernels", 0)
            e2e_malloc = pack("e2e_malloc", "same call with MP3D_BATCH_OUT_MALLOC (drop-in: one malloc()ed buffer per stream, "
                                            "freed by the caller inside the timed loop's untimed tail)", 2*total_samples)

    if rank == 0:
        peak, peak_kind = measured_peaks()
        ent_bps = st["input_bytes"]/max(1, samples_per_step) + 4.0
        bps = {"transform": TRANSFORM_BYTES_PER_SAMPLE, "entropy": ent_bps}
        dom = max(("entropy", "transform"), key=lambda k: kern.get(k, 0.0))
        l3_samples = samples_per_step     # Layer I/II blocks do not pass through these two kernels
        if st["n_granule_channels"]:
            l3_samples = min(samples_per_step, 576*int(st["n_granule_channels"]))
        achieved = bps[dom]*l3_samples/(kern[dom]*1e-3)/1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp) and args.config == corpus.DEFAULT_CONFIG:
            try:
                traffic = json.load(open(tp)).get(dom)
            except Exception:
                traffic = None
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        fp32_peak = FP32_LANES*2*sm_mhz*1e6/1e12
        fp32_ach = TRANSFORM_FLOPS_PER_SAMPLE*l3_samples/(kern["transform"]*1e-3)/1e12 if kern.get("transform") else None
        hbm_frac_tr = TRANSFORM_BYTES_PER_SAMPLE*l3_samples/(kern["transform"]*1e-3)/1e9/peak if kern.get("transform") else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max/args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": data_string(args),
                "value_scope": "kernels_resident: every GPU kernel of the job (side-info parse, lane order, entropy, stereo, "
                               "tile meta, transform, Layer I/II); bitstreams + frame records already in HBM",
                "config": config_dict(args, world),
                "workload_stats": {"frames_per_gpu": int(st["n_frames"]), "granule_channels_per_gpu": int(st["n_granule_channels"]),
                                   "samples_per_step_per_gpu": samples_per_step, "input_bytes_per_gpu": int(st["input_bytes"]),
                                   "step_working_set_mb": (st["input_bytes"] + 4*576*st["n_granule_channels"] + 2*samples_per_step)/1e6,
                                   "plan_host_parse_ms": plan_parse_ms, "plan_h2d_ms": plan_h2d_ms, "plan_h2d_bytes": h2d},
                "gpu_launches": (int(st["kernels_per_run"]) + int(st["kernels_prep"]))*args.steps,
                "kernels_ms": kern,
                "roofline": {"bound": "hbm", "kernel": "k_l3_" + dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved/peak, "traffic": traffic, "peak_kind": peak_kind,
                             "bytes_per_sample": bps[dom],
                             "note": "HBM is the lower of the two nominal ceilings (6 B/sample vs 56 flop/sample), so it stays "
                                     "`bound`; the kernel itself is issue/FP32/shared-memory limited (profiles/), see fp32",
                             "fp32": {"kernel": "k_l3_transform", "achieved": fp32_ach, "peak": fp32_peak, "unit": "TFLOP/s",
                                      "frac": (fp32_ach/fp32_peak) if fp32_ach else None,
                                      "flops_per_sample": TRANSFORM_FLOPS_PER_SAMPLE,
                                      "peak_how": "148 SM x 128 FFMA lanes x 2 x sampled SM clock"},
                             "transform_hbm_frac": hbm_frac_tr},
                "clocks": clocks}
        if e2e is not None:
            line["e2e"] = e2e
            line["scope_ii"] = scope_ii
            line["e2e_malloc"] = e2e_malloc
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle import refshim
                refshim.build()
                ref = refshim.Ref()
                # bounded sample: a stride through the batch worth ~10-20 s of one core per thread at most
                n_cpu = min(len(streams), max(cores*8, 64))
                pick = [streams[(i*len(streams))//n_cpu] for i in range(n_cpu)]
                ref.bench_decode(pick[:cores], cores)
                s_cpu, t_cpu = ref.bench_decode(pick, cores)
                line["cpu_baseline"] = {"value": s_cpu/t_cpu, "unit": UNIT, "cores": cores, "kind": "reference",
                                        "sample": "%d of the %d streams (even stride), 1 pass, unmodified minimp3 (SSE2, gcc -O2), pthreads"
                                                  % (n_cpu, len(streams))}
            except Exception as e:  # the checker being absent must not hide the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
