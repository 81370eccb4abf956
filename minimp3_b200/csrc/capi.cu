/*
 * capi.cu -- the C ABI of libminimp3_b200.so (include/minimp3_b200.h): batch plans, the drop-in
 * mp3dec_* entry points, buffer management and stream plumbing around the three kernels.
 *
 * There is deliberately no CPU decode path here: when CUDA is unavailable every entry point that
 * would produce PCM returns MP3D_E_CUDA (or 0 samples for the frame API, after printing why).
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/minimp3_b200.h"
#include "host_parse.h"
#include "kernels.h"
#include "l12_decode.cuh"
#include "mp3_hdr.cuh"

using namespace mp3b;

namespace {

double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

/* Persistent worker pool for the host stage (frame walk, rebasing, output copies). */
class HostPool
{
public:
    /* the pool of the calling thread: the process-wide one (all cores), or, inside a multi-device call, the
     * device thread's share of the cores (set_current) so that the devices' host stages run side by side */
    static HostPool &get()
    {
        if (current_) return *current_;
        static HostPool p((int)std::thread::hardware_concurrency());
        return p;
    }
    static void set_current(HostPool *p) { current_ = p; }
    explicit HostPool(int n)
    {
        if (n <= 0) n = 1;
        for (int t = 0; t < n - 1; t++) workers_.emplace_back([this, t] { loop(t); });
    }
    ~HostPool()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            generation_++;
        }
        cv_.notify_all();
        for (auto &w : workers_) w.join();
    }
    /* runs f(i) for i in [0, n) on up to `threads` workers (0 = all) plus the calling thread */
    void run(int n, int threads, const std::function<void(int)> &f)
    {
        if (n <= 0) return;
        int want = threads <= 0 ? (int)workers_.size() + 1 : threads;
        want = std::min(want, n);
        if (want <= 1 || workers_.empty())
        {
            for (int i = 0; i < n; i++) f(i);
            return;
        }
        std::unique_lock<std::mutex> call_lock(call_mu_);     /* one parallel region at a time */
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &f; n_ = n; next_.store(0); active_ = 0; helpers_ = std::min(want - 1, (int)workers_.size());
            generation_++;
        }
        cv_.notify_all();
        for (;;)
        {
            const int i = next_.fetch_add(1);
            if (i >= n) break;
            f(i);
        }
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [&] { return active_ == 0 && started_ == helpers_; });
        fn_ = nullptr;
        started_ = 0;
    }

private:
    static thread_local HostPool *current_;
    void loop(int id)
    {
        uint64_t seen = 0;
        for (;;)
        {
            const std::function<void(int)> *fn;
            int n;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return generation_ != seen; });
                seen = generation_;
                if (stop_) return;
                if (id >= helpers_ || !fn_) continue;
                fn = fn_; n = n_;
                active_++; started_++;
            }
            for (;;)
            {
                const int i = next_.fetch_add(1);
                if (i >= n) break;
                (*fn)(i);
            }
            {
                std::lock_guard<std::mutex> lk(mu_);
                active_--;
            }
            done_cv_.notify_all();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_, call_mu_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int)> *fn_ = nullptr;
    int n_ = 0, active_ = 0, started_ = 0, helpers_ = 0;
    std::atomic<int> next_{0};
    uint64_t generation_ = 0;
    bool stop_ = false;
};

thread_local HostPool *HostPool::current_ = nullptr;

template <class F>
void parallel_for(int n, int threads, F f)
{
    HostPool::get().run(n, threads, std::function<void(int)>(f));
}

struct DevBuf
{
    void *p = nullptr;
    size_t cap = 0;
    bool ensure(size_t n)
    {
        if (n <= cap) return true;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = n + n/8 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) { p = nullptr; return false; }
        cap = want;
        return true;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PinBuf
{
    void *p = nullptr;
    size_t cap = 0;
    bool ensure(size_t n)
    {
        if (n <= cap) return true;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        const size_t want = n + n/8 + 256;
        if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) { p = nullptr; return false; }
        cap = want;
        return true;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

/* all device + pinned memory of one in-flight batch; recycled between calls */
struct BufferSet
{
    PinBuf h_md, h_gcs, h_frames, h_tiles, h_ist, h_order, h_gcbase, h_pcm, h_l12, h_tiles12;
    DevBuf d_md, d_gcs, d_frames, d_tiles, d_tile_recs, d_tile_jobs, d_counter, d_istlist, d_order, d_order_sorted, d_order_scratch, d_gcbase, d_xr, d_nzl, d_istpos, d_pcm, d_l12, d_tiles12, d_sub12, d_l12info;
    DevBuf d_tap_q, d_tap_scf, d_tap_sub, d_phase_clk;
    /* GPU frame scan (k_scan.cu) */
    PinBuf h_scan, h_sum, h_bases, h_slotmap;
    DevBuf d_raw, d_scan, d_scan_frames, d_sum, d_bases, d_slotmap;
    /* what this set has served so far: route (0 host walker, 1 GPU frame scan) and the largest input of a plan.  A call
     * cuts its streams into chunks of different sizes; handing every plan the set that last held a plan of its kind
     * and size keeps repeated calls from re-growing (cudaFree + cudaMalloc / cudaHostAlloc) a too-small set each time */
    int route = -1;
    size_t served_bytes = 0;
    void release()
    {
        h_md.release(); h_gcs.release(); h_frames.release(); h_tiles.release(); h_ist.release(); h_order.release(); h_gcbase.release(); h_pcm.release(); h_l12.release(); h_tiles12.release();
        d_l12.release(); d_tiles12.release(); d_sub12.release(); d_l12info.release();
        d_md.release(); d_gcs.release(); d_frames.release(); d_tiles.release(); d_tile_recs.release(); d_tile_jobs.release(); d_counter.release(); d_istlist.release(); d_order.release(); d_order_sorted.release(); d_order_scratch.release(); d_gcbase.release(); d_xr.release(); d_nzl.release();
        h_scan.release(); h_sum.release(); h_bases.release(); h_slotmap.release();
        d_raw.release(); d_scan.release(); d_scan_frames.release(); d_sum.release(); d_bases.release(); d_slotmap.release();
        d_istpos.release(); d_pcm.release(); d_tap_q.release(); d_tap_scf.release(); d_tap_sub.release(); d_phase_clk.release();
    }
};

struct DeviceCtx
{
    int device = -1;
    bool ready = false;
    DeviceTables tables{};
    cudaStream_t stream = nullptr;      /* H2D + kernels */
    cudaStream_t stream_d2h = nullptr;  /* PCM read-back, overlaps the next chunk's kernels */
    cudaStream_t stream_h2d = nullptr;  /* chunk-pipelined calls: inputs of chunk c+1 go up while chunk c computes */
    std::mutex mu;
    std::vector<BufferSet *> free_sets;
    /* sets whose PCM (pinned slab / device buffer) was handed to the caller by a MP3D_BATCH_OUT_PINNED / _DEVICE
     * call: nobody reuses them -- not the library's internal batch calls either -- until the NEXT call of that kind
     * on this device (or mp3dec_batch_release_outputs) */
    std::vector<BufferSet *> exported_sets;
    /* single-frame API scratch */
    BufferSet frame_set;
    DevBuf d_state_in, d_state_out;
    PinBuf h_state;
};

std::mutex g_ctx_mu;
DeviceCtx *g_ctx[64];
cudaEvent_t g_dbg_base = nullptr;     /* MP3B_DEBUG_TIMING: start of the current call on the read-back stream */

DeviceCtx *get_ctx(int device)
{
    if (device < 0 || device >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(g_ctx_mu);
    if (g_ctx[device] && g_ctx[device]->ready) return g_ctx[device];
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device >= n)
    {
        fprintf(stderr, "minimp3_b200: no usable CUDA device %d (this library has no CPU decode path)\n", device);
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    DeviceCtx *c = new DeviceCtx();
    c->device = device;
    if (tables_upload(&c->tables) != 0 || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->stream_d2h, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->stream_h2d, cudaStreamNonBlocking) != cudaSuccess)
    {
        fprintf(stderr, "minimp3_b200: CUDA initialisation failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        delete c;
        return nullptr;
    }
    c->ready = true;
    g_ctx[device] = c;
    return c;
}

BufferSet *acquire_set(DeviceCtx *c, int route, size_t input_bytes)
{
    std::lock_guard<std::mutex> lock(c->mu);
    /* best fit: same route first; the smallest set that has held at least this much, else the largest one */
    int best = -1;
    auto better = [&](const BufferSet *a, const BufferSet *b) {   /* a preferable to b */
        const bool ra = a->route == route, rb = b->route == route;
        if (ra != rb) return ra;
        const bool fa = a->served_bytes >= input_bytes, fb = b->served_bytes >= input_bytes;
        if (fa != fb) return fa;
        return fa ? a->served_bytes < b->served_bytes : a->served_bytes > b->served_bytes;
    };
    for (int i = 0; i < (int)c->free_sets.size(); i++)
        if (best < 0 || better(c->free_sets[i], c->free_sets[best])) best = i;
    BufferSet *s;
    if (best >= 0)
    {
        s = c->free_sets[best];
        c->free_sets.erase(c->free_sets.begin() + best);
    } else
        s = new BufferSet();
    if (s->route != route) { s->route = route; s->served_bytes = 0; }
    if (s->served_bytes < input_bytes) s->served_bytes = input_bytes;
    return s;
}

void release_set(DeviceCtx *c, BufferSet *s)
{
    std::lock_guard<std::mutex> lock(c->mu);
    if (c->free_sets.size() < 48) c->free_sets.push_back(s);
    else { s->release(); delete s; }
}

void release_exported(DeviceCtx *c)
{
    std::vector<BufferSet *> sets;
    {
        std::lock_guard<std::mutex> lock(c->mu);
        sets.swap(c->exported_sets);
    }
    for (BufferSet *s : sets) release_set(c, s);
}

size_t round_up(size_t v, size_t a) { return (v + a - 1)/a*a; }

}  // namespace

/* fewer granule-channels or streams than this: the per-stream order of the host walker is used as it is */
static const uint64_t kMinOrderSort = 8192;

struct mp3dec_batch_plan
{
    DeviceCtx *ctx = nullptr;
    BufferSet *bufs = nullptr;
    mp3dec_batch_opts_t opts{};
    cudaStream_t stream = nullptr;
    int n_streams = 0;
    std::vector<StreamPlan> plans;
    std::vector<uint64_t> md_base, gc_base, fr_base, tile_base, pcm_base, ist_base, order_base, l12_base, blk_base, tile12_base;
    uint64_t n_l12 = 0, n_blk = 0, n_tiles12 = 0;
    uint64_t md_bytes = 0, n_gc = 0, n_tiles = 0, n_samples = 0, n_ist = 0, n_order = 0, n_frames = 0, input_bytes = 0;
    size_t sample_size = 2;
    bool taps = false;
    cudaEvent_t ev_kernels = nullptr, ev_d2h = nullptr, ev_h2d = nullptr;
    cudaEvent_t ev_dbg0 = nullptr;        /* MP3B_DEBUG_TIMING only */
    cudaStream_t scan_stream = nullptr;   /* GPU frame scan route: upload + scan + record kernels of this plan */
    std::vector<int> out_index;    /* stream i of the plan is stream out_index[i] of the caller (empty: identity) */
    bool export_on_destroy = false; /* one-call entry point with pinned / device output: the caller holds pointers into bufs */
    bool order_sorted = false;     /* the entropy kernel reads d_order_sorted (order across streams, k_order.cu) */
    bool sort_order = false;
    size_t order_scratch = 0;
    uint32_t max_slots = 0;
    uint64_t n_frame_recs = 0;
    bool d2h_started = false;
    mp3dec_batch_stats_t stats{};
};

extern "C" {

const char *mp3dec_b200_version(void) { return "minimp3_b200 0.1 (sm_100a)"; }

void mp3dec_batch_shard(const size_t *sizes, int n_streams, int n_shards, int *shard_of)
{
    if (n_shards < 1) n_shards = 1;
    std::vector<int> order(n_streams);
    for (int i = 0; i < n_streams; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return sizes[a] > sizes[b]; });
    std::vector<uint64_t> load(n_shards, 0);
    for (int k = 0; k < n_streams; k++)
    {
        int best = 0;
        for (int s = 1; s < n_shards; s++)
            if (load[s] < load[best]) best = s;
        shard_of[order[k]] = best;
        load[best] += sizes[order[k]];
    }
}

static mp3dec_batch_plan_t *plan_create_impl(const mp3dec_batch_stream_t *streams, int n_streams,
                                             const mp3dec_batch_opts_t *opts_in, int *err, bool async);

/* GPU preparation of a plan, behind the H2D copies: side info -> descriptors (k_l3_side_info) and the lane order of
 * the entropy kernel across streams (k_order.cu).  Runs once at plan creation; stage 4 of
 * mp3dec_batch_plan_run_stage repeats it so that a measurement can include every kernel of the job. */
static int plan_prep(mp3dec_batch_plan_t *P)
{
    BufferSet &B = *P->bufs;
    if (P->n_frame_recs && launch_side_info((const FrameRec *)B.d_frames.p, (uint32_t)P->n_frame_recs, (GcDesc *)B.d_gcs.p, P->stream) != cudaSuccess)
        return MP3D_E_CUDA;
    /* the walker ordered every stream's granule-channels by size; with enough of them the GPU redoes it across the
     * streams of a window, which is what fills the entropy kernel's warps evenly when streams are alike */
    if (P->sort_order)
    {
        if (launch_order_sort((const GcDesc *)B.d_gcs.p, (const uint32_t *)B.d_order.p, (uint32_t)P->n_order, (uint32_t *)B.d_order_sorted.p,
                              (const uint32_t *)B.d_gcbase.p, (uint32_t)P->n_streams, P->max_slots, B.d_order_scratch.p, P->order_scratch,
                              P->stream) != cudaSuccess)
            return MP3D_E_CUDA;
        P->order_sorted = true;
    }
    return 0;
}

static mp3dec_batch_plan_t *plan_create_scan(const mp3dec_batch_stream_t *streams, int n_streams, const mp3dec_batch_opts_t *opts_in,
                                             int *err, std::vector<int> *rest);

mp3dec_batch_plan_t *mp3dec_batch_plan_create(const mp3dec_batch_stream_t *streams, int n_streams,
                                              const mp3dec_batch_opts_t *opts_in, int *err)
{
    int dummy;
    if (!err) err = &dummy;
    if (opts_in && opts_in->gpu_frame_scan > 0 && streams && n_streams > 0)
    {
        /* a plan is one set of buffers: the scan route is taken when it can take every stream */
        bool ok = true;
        for (int i = 0; i < n_streams; i++) ok &= !((!streams[i].buf && streams[i].size) || streams[i].size == (size_t)-1);
        if (ok)
        {
            std::vector<int> rest;
            mp3dec_batch_plan_t *P = plan_create_scan(streams, n_streams, opts_in, err, &rest);
            if (P && rest.empty() && cudaStreamSynchronize(P->stream) == cudaSuccess) return P;
            if (P) { cudaStreamSynchronize(P->stream); mp3dec_batch_plan_destroy(P); }
        }
    }
    return plan_create_impl(streams, n_streams, opts_in, err, false);
}

static mp3dec_batch_plan_t *plan_create_impl(const mp3dec_batch_stream_t *streams, int n_streams,
                                             const mp3dec_batch_opts_t *opts_in, int *err, bool async)
{
    int dummy;
    if (!err) err = &dummy;
    *err = 0;
    if (!streams || n_streams < 0) { *err = MP3D_E_PARAM; return nullptr; }
    mp3dec_batch_opts_t opts;
    memset(&opts, 0, sizeof(opts));
    if (opts_in) opts = *opts_in;
    for (int i = 0; i < n_streams; i++)
        if ((!streams[i].buf && streams[i].size) || streams[i].size == (size_t)-1) { *err = MP3D_E_PARAM; return nullptr; }
    DeviceCtx *ctx = get_ctx(opts.device);
    if (!ctx) { *err = MP3D_E_CUDA; return nullptr; }
    cudaSetDevice(ctx->device);

    mp3dec_batch_plan *P = new mp3dec_batch_plan();
    P->ctx = ctx;
    P->opts = opts;
    P->stream = opts.cuda_stream ? (cudaStream_t)opts.cuda_stream : ctx->stream;
    P->n_streams = n_streams;
    P->sample_size = opts.float_output ? 4 : 2;
    {
        size_t in_bytes = 0;
        for (int i = 0; i < n_streams; i++) in_bytes += streams[i].size;
        P->bufs = acquire_set(ctx, 0, in_bytes);
    }
    BufferSet &B = *P->bufs;

    const double t0 = now_ms();
    /* ---- host stage 1: main-data staging offsets, then one walker per stream ---- */
    P->md_base.resize(n_streams + 1);
    uint64_t off = 0;
    for (int i = 0; i < n_streams; i++)
    {
        P->md_base[i] = off;
        off += round_up(streams[i].size + 16, 16);
    }
    P->md_base[n_streams] = off;
    P->md_bytes = off + 4096;      /* guard so that run-away Huffman reads stay inside the buffer */
    if (!B.h_md.ensure(P->md_bytes)) { *err = MP3D_E_MEMORY; mp3dec_batch_plan_destroy(P); return nullptr; }
    P->plans.resize(n_streams);
    WalkOptions wopt;
    wopt.raw_frames = opts.raw_frames != 0;
    wopt.allow_mono_stereo = opts.allow_mono_stereo_transition != 0;
    wopt.device_side_info = true;        /* descriptors are produced on the GPU from FrameRec records (k_sideinfo.cu) */
    uint8_t *h_md = (uint8_t *)B.h_md.p;
    parallel_for(n_streams, opts.host_threads, [&](int i) {
        uint8_t *dst = h_md + P->md_base[i];
        walk_stream(streams[i].buf, streams[i].size, dst, P->md_base[i]*8u, wopt, &P->plans[i]);
        /* zero the slack so the device never sees stale bytes of an earlier batch */
        const uint64_t used = P->plans[i].main_data_bytes, room = P->md_base[i + 1] - P->md_base[i];
        memset(dst + used, 0, (size_t)(room - used));
    });
    memset(h_md + off, 0, 4096);

    /* ---- host stage 2: prefix sums and rebasing into the shared tables ---- */
    P->gc_base.resize(n_streams + 1); P->tile_base.resize(n_streams + 1);
    P->pcm_base.resize(n_streams + 1); P->ist_base.resize(n_streams + 1); P->order_base.resize(n_streams + 1);
    P->l12_base.resize(n_streams + 1); P->blk_base.resize(n_streams + 1); P->tile12_base.resize(n_streams + 1);
    uint64_t gcs = 0, tiles = 0, samples = 0, ists = 0, orders = 0, l12s = 0, blks = 0, tiles12 = 0, frs = 0;
    P->fr_base.resize(n_streams + 1);
    for (int i = 0; i < n_streams; i++)
    {
        const StreamPlan &sp = P->plans[i];
        P->gc_base[i] = gcs; P->tile_base[i] = tiles; P->pcm_base[i] = samples; P->ist_base[i] = ists; P->order_base[i] = orders;
        orders += sp.order.size();
        P->l12_base[i] = l12s; P->blk_base[i] = blks; P->tile12_base[i] = tiles12;
        l12s += sp.l12_frames.size(); blks += sp.n_blockch; tiles12 += sp.tiles12.size();
        gcs += round_up((size_t)sp.n_gc, 2);
        P->fr_base[i] = frs;
        frs += sp.frame_recs.size();
        tiles += sp.tiles.size();
        samples += sp.samples_decoded;
        ists += sp.istereo_gcs.size();
        P->n_frames += sp.frames_decoded;
        P->input_bytes += sp.main_data_bytes;
    }
    P->gc_base[n_streams] = gcs; P->tile_base[n_streams] = tiles; P->pcm_base[n_streams] = samples; P->ist_base[n_streams] = ists;
    P->n_gc = gcs; P->n_tiles = tiles; P->n_samples = samples; P->n_ist = ists; P->n_order = orders;
    P->n_l12 = l12s; P->n_blk = blks; P->n_tiles12 = tiles12;
    if (gcs >= (1ull << 31)) { *err = MP3D_E_PARAM; mp3dec_batch_plan_destroy(P); return nullptr; }

    bool ok = B.h_frames.ensure((size_t)frs*sizeof(FrameRec) + 64) && B.h_tiles.ensure((size_t)tiles*sizeof(TileDesc) + 64) &&
              B.h_ist.ensure((size_t)ists*4 + 64) && B.h_order.ensure((size_t)orders*4 + 64) &&
              B.h_l12.ensure((size_t)l12s*sizeof(L12FrameDesc) + 64) && B.h_tiles12.ensure((size_t)tiles12*sizeof(TileDesc) + 64);
    if (!ok) { *err = MP3D_E_MEMORY; mp3dec_batch_plan_destroy(P); return nullptr; }
    FrameRec *h_frames = (FrameRec *)B.h_frames.p;
    TileDesc *h_tiles = (TileDesc *)B.h_tiles.p;
    uint32_t *h_ist = (uint32_t *)B.h_ist.p;
    uint32_t *h_order = (uint32_t *)B.h_order.p;
    L12FrameDesc *h_l12 = (L12FrameDesc *)B.h_l12.p;
    TileDesc *h_tiles12 = (TileDesc *)B.h_tiles12.p;
    parallel_for(n_streams, opts.host_threads, [&](int i) {
        StreamPlan &sp = P->plans[i];
        const uint64_t gb = P->gc_base[i];
        for (size_t k = 0; k < sp.frame_recs.size(); k++)
        {
            FrameRec r = sp.frame_recs[k];
            r.gc_first += (uint32_t)gb;
            h_frames[P->fr_base[i] + k] = r;
        }
        for (size_t k = 0; k < sp.tiles.size(); k++)
        {
            TileDesc t = sp.tiles[k];
            t.gc_first += (uint32_t)gb;
            t.pcm_off += P->pcm_base[i];
            for (int c = 0; c < 2; c++)
                for (int h = 0; h < 2; h++)
                    if (t.halo[c][h] >= 0) t.halo[c][h] += (int32_t)gb;
            h_tiles[P->tile_base[i] + k] = t;
        }
        for (size_t k = 0; k < sp.istereo_gcs.size(); k++) h_ist[P->ist_base[i] + k] = sp.istereo_gcs[k] + (uint32_t)gb;
        for (size_t k = 0; k < sp.order.size(); k++) h_order[P->order_base[i] + k] = sp.order[k] + (uint32_t)gb;
        std::vector<uint32_t>().swap(sp.order);
        const uint64_t bb = P->blk_base[i];
        for (size_t k = 0; k < sp.l12_frames.size(); k++)
        {
            L12FrameDesc f = sp.l12_frames[k];
            f.block_first += (uint32_t)bb;
            h_l12[P->l12_base[i] + k] = f;
        }
        for (size_t k = 0; k < sp.tiles12.size(); k++)
        {
            TileDesc t = sp.tiles12[k];
            t.gc_first += (uint32_t)bb;
            t.pcm_off += P->pcm_base[i];
            for (int c = 0; c < 2; c++)
                for (int h = 0; h < 2; h++)
                    if (t.halo[c][h] >= 0) t.halo[c][h] += (int32_t)bb;
            h_tiles12[P->tile12_base[i] + k] = t;
        }
        std::vector<L12FrameDesc>().swap(sp.l12_frames);
        std::vector<TileDesc>().swap(sp.tiles12);
        /* descriptor vectors are no longer needed on the host */
        std::vector<GcDesc>().swap(sp.gcs);
        std::vector<FrameRec>().swap(sp.frame_recs);
        std::vector<uint32_t>().swap(sp.gc_key);
        std::vector<TileDesc>().swap(sp.tiles);
        std::vector<GranuleRef>().swap(sp.granules);
        std::vector<uint32_t>().swap(sp.istereo_gcs);
    });
    const double t1 = now_ms();

    /* ---- device buffers + H2D ---- */
    const bool sort_order = frs && orders >= kMinOrderSort && n_streams >= 2;
    const size_t order_scratch = sort_order ? order_sort_scratch_bytes((uint32_t)orders, (uint32_t)n_streams) : 0;
    uint32_t max_slots = 0;
    if (sort_order)
    {
        if (!B.h_gcbase.ensure(((size_t)n_streams + 1)*4 + 64)) { *err = MP3D_E_MEMORY; mp3dec_batch_plan_destroy(P); return nullptr; }
        uint32_t *gb = (uint32_t *)B.h_gcbase.p;
        for (int i = 0; i <= n_streams; i++) gb[i] = (uint32_t)P->gc_base[i];
        for (int i = 0; i < n_streams; i++) max_slots = std::max(max_slots, gb[i + 1] - gb[i]);
    }
    ok = B.d_md.ensure(P->md_bytes) && B.d_gcs.ensure((size_t)gcs*sizeof(GcDesc) + 64) && B.d_frames.ensure((size_t)frs*sizeof(FrameRec) + 64) &&
         B.d_tiles.ensure((size_t)tiles*sizeof(TileDesc) + 64) && B.d_tile_recs.ensure((size_t)tiles*sizeof(TileRec) + 256) && B.d_tile_jobs.ensure((size_t)tiles*MP3B_TILE_JOBS*2 + 256) && B.d_counter.ensure(256) &&
         B.d_istlist.ensure((size_t)ists*4 + 64) &&
         B.d_order.ensure((size_t)orders*4 + 64) &&
         (!sort_order || (B.d_order_sorted.ensure((size_t)orders*4 + 64) && B.d_order_scratch.ensure(order_scratch) &&
                          B.d_gcbase.ensure(((size_t)n_streams + 1)*4 + 64))) &&
         B.d_l12.ensure((size_t)l12s*sizeof(L12FrameDesc) + 64) &&
         B.d_tiles12.ensure((size_t)tiles12*sizeof(TileDesc) + 64) && B.d_sub12.ensure((size_t)blks*384*sizeof(float) + 64) &&
         B.d_l12info.ensure((size_t)l12s*sizeof(L12Compact) + 64) &&
         B.d_xr.ensure((size_t)gcs*576*sizeof(float) + 64) && B.d_nzl.ensure((size_t)gcs*2 + 64) &&
         B.d_istpos.ensure((size_t)gcs*40 + 64) && B.d_pcm.ensure((size_t)samples*P->sample_size + 64);
    if (!ok) { *err = MP3D_E_MEMORY; mp3dec_batch_plan_destroy(P); return nullptr; }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (!async) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, P->stream); }
    /* in a chunk-pipelined call the inputs travel on their own stream, so that they overlap the kernels of the
     * chunk before; the plan's stream picks up behind them */
    cudaStream_t up = (async && !opts.cuda_stream) ? ctx->stream_h2d : P->stream;
    cudaMemcpyAsync(B.d_md.p, B.h_md.p, P->md_bytes, cudaMemcpyHostToDevice, up);
    if (gcs) cudaMemsetAsync(B.d_gcs.p, 0, (size_t)gcs*sizeof(GcDesc), up);   /* padding slots are never written: keep them defined (the lane-order heuristic looks at them) */
    if (frs) cudaMemcpyAsync(B.d_frames.p, B.h_frames.p, (size_t)frs*sizeof(FrameRec), cudaMemcpyHostToDevice, up);
    if (tiles) cudaMemcpyAsync(B.d_tiles.p, B.h_tiles.p, (size_t)tiles*sizeof(TileDesc), cudaMemcpyHostToDevice, up);
    if (ists) cudaMemcpyAsync(B.d_istlist.p, B.h_ist.p, (size_t)ists*4, cudaMemcpyHostToDevice, up);
    if (orders) cudaMemcpyAsync(B.d_order.p, B.h_order.p, (size_t)orders*4, cudaMemcpyHostToDevice, up);
    if (sort_order) cudaMemcpyAsync(B.d_gcbase.p, B.h_gcbase.p, ((size_t)n_streams + 1)*4, cudaMemcpyHostToDevice, up);
    if (l12s) cudaMemcpyAsync(B.d_l12.p, B.h_l12.p, (size_t)l12s*sizeof(L12FrameDesc), cudaMemcpyHostToDevice, up);
    if (tiles12) cudaMemcpyAsync(B.d_tiles12.p, B.h_tiles12.p, (size_t)tiles12*sizeof(TileDesc), cudaMemcpyHostToDevice, up);
    if (up != P->stream)
    {
        cudaEventCreateWithFlags(&P->ev_h2d, cudaEventDisableTiming);
        cudaEventRecord(P->ev_h2d, up);
        cudaStreamWaitEvent(P->stream, P->ev_h2d, 0);
    }
    P->sort_order = sort_order; P->order_scratch = order_scratch; P->max_slots = max_slots; P->n_frame_recs = frs;
    if (plan_prep(P) != 0)
    {
        *err = MP3D_E_CUDA;
        mp3dec_batch_plan_destroy(P);
        return nullptr;
    }
    float h2d = 0.f;
    if (!async)
    {
        cudaEventRecord(e1, P->stream);
        if (cudaStreamSynchronize(P->stream) != cudaSuccess)
        {
            fprintf(stderr, "minimp3_b200: H2D failed: %s\n", cudaGetErrorString(cudaGetLastError()));
            *err = MP3D_E_CUDA;
            cudaEventDestroy(e0); cudaEventDestroy(e1);
            mp3dec_batch_plan_destroy(P);
            return nullptr;
        }
        cudaEventElapsedTime(&h2d, e0, e1);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }

    P->stats.input_bytes = P->input_bytes;
    P->stats.samples = samples;
    P->stats.n_granule_channels = gcs;
    P->stats.n_tiles = tiles;
    P->stats.n_istereo_granules = ists;
    P->stats.n_frames = P->n_frames;
    P->stats.h2d_bytes = P->md_bytes + frs*sizeof(FrameRec) + tiles*sizeof(TileDesc) + ists*4 + orders*4 +
                         l12s*sizeof(L12FrameDesc) + tiles12*sizeof(TileDesc);
    P->stats.kernels_per_run = (orders ? 1 : 0) + (ists ? 1 : 0) + (tiles ? 2 : 0) + (l12s ? 2 : 0) + (tiles12 ? 1 : 0);
    P->stats.kernels_prep = (frs ? 1 : 0) + (sort_order ? 2 : 0);    /* own kernels; the radix sort behind them is CUB's */
    P->stats.host_parse_ms = t1 - t0;
    P->stats.h2d_ms = h2d;
    return P;
}

/* ---- plan creation through the GPU frame scan (k_scan.cu) ------------------------------------------------------
 * The host finds where every stream's audio starts (tags, first frame, VBR tag: open_stream_head / find_frame) and
 * copies the streams into pinned staging; the per-frame work -- header chain, side-info checks, reservoir
 * accounting, payload concatenation, frame records, tiles -- runs on the device.  Streams the scan cannot take
 * (Layer I/II, free format, anything irregular: see k_scan.cu) are returned in *rest for the host walker.  The plan
 * covers the streams it took plus the ones with nothing to decode. */
struct ScanHead
{
    const uint8_t *p = nullptr;    /* first audio frame */
    size_t size = 0;
    StreamHead h;
    int kind = 0;                  /* 0 nothing to decode, 1 scan candidate, 2 host walker */
    uint32_t slot_cap = 0;
};

/* a scan in flight: everything up to the scan kernel is queued, the summaries are on their way back */
struct ScanJob
{
    mp3dec_batch_plan *P = nullptr;
    mp3dec_batch_opts_t opts;
    int n_streams = 0, nc = 0;
    std::vector<ScanHead> heads;
    std::vector<int> cand, rest;
    uint64_t raw_bytes = 0, n_slots = 0;
    cudaStream_t up = nullptr;
    cudaEvent_t ev_sum = nullptr;
    double t0 = 0, t_heads = 0, t_issue = 0;
};

static void scan_job_free(ScanJob *J)
{
    if (!J) return;
    if (J->ev_sum) cudaEventDestroy(J->ev_sum);
    delete J;
}

/* first half: stream heads on the host, staging copy, upload, scan kernel, summaries on their way back (asynchronous) */
static ScanJob *scan_begin(const mp3dec_batch_stream_t *streams, int n_streams, const mp3dec_batch_opts_t *opts_in, int *err)
{
    *err = 0;
    mp3dec_batch_opts_t opts;
    memset(&opts, 0, sizeof(opts));
    if (opts_in) opts = *opts_in;
    DeviceCtx *ctx = get_ctx(opts.device);
    if (!ctx) { *err = MP3D_E_CUDA; return nullptr; }
    cudaSetDevice(ctx->device);
    const double t0 = now_ms();

    ScanJob *J = new ScanJob();
    J->opts = opts;
    J->n_streams = n_streams;
    J->t0 = t0;
    J->heads.resize(n_streams);
    std::vector<ScanHead> &heads = J->heads;
    std::vector<int> *rest = &J->rest;
    parallel_for(n_streams, opts.host_threads, [&](int i) {
        ScanHead &H = heads[i];
        const uint8_t *buf = streams[i].buf;
        size_t size = streams[i].size;
        if (size >= (1u << 31)) { H.kind = 2; return; }
        if (!opts.raw_frames)
        {
            if (!open_stream_head(&buf, &size, &H.h)) return;
        } else
        {
            int free_format_bytes = 0, frame_size = 0;
            const int junk = find_frame(buf, (int)size, &free_format_bytes, &frame_size);
            if (!frame_size) return;
            buf += junk;
            size -= (size_t)junk;
        }
        H.p = buf;
        H.size = size;
        if (size < 4 || hd_layer(buf) != 3 || hd_is_free_format(buf)) { H.kind = 2; return; }
        /* shortest frame this stream can hold: lowest bit rate of its version / sample rate, no padding */
        uint8_t probe[4] = { buf[0], buf[1], (uint8_t)((buf[2] & 0x0C) | 0x10), buf[3] };
        const int min_fs = std::max(hd_frame_bytes(probe, 0), 24);
        H.slot_cap = (uint32_t)(size/(size_t)min_fs) + 2;
        H.kind = 1;
    });
    J->t_heads = now_ms();
    /* candidates: staging offsets and slot groups */
    std::vector<int> &cand = J->cand;
    std::vector<uint64_t> raw_base;
    uint64_t raw_bytes = 0, n_slots = 0;
    for (int i = 0; i < n_streams; i++)
    {
        if (heads[i].kind == 2) rest->push_back(i);
        if (heads[i].kind != 1) continue;
        cand.push_back(i);
        raw_base.push_back(raw_bytes);
        raw_bytes += round_up(heads[i].size + 16, 16);
        n_slots += heads[i].slot_cap;
    }
    const int nc = (int)cand.size();
    /* direct upload: every candidate inside the caller's page-locked region, and the span they cover not much
     * larger than their bytes (it is uploaded whole, gaps included) */
    bool direct = false;
    const uint8_t *slab_lo = nullptr;
    if (nc && opts.pinned_inputs && opts.pinned_inputs_bytes)
    {
        const uint8_t *lo = heads[cand[0]].p, *hi = lo;
        uint64_t sum = 0;
        for (int k : cand)
        {
            lo = std::min(lo, heads[k].p);
            hi = std::max(hi, heads[k].p + heads[k].size);
            sum += heads[k].size;
        }
        const uint8_t *r0 = (const uint8_t *)opts.pinned_inputs, *r1 = r0 + opts.pinned_inputs_bytes;
        if (lo >= r0 && hi <= r1 && (uint64_t)(hi - lo) <= sum + sum/4 + 65536)
        {
            direct = true;
            slab_lo = lo;
            raw_bytes = (uint64_t)(hi - lo);
        }
    }
    J->nc = nc; J->raw_bytes = raw_bytes; J->n_slots = n_slots;
    if (n_slots >= (1ull << 31)) { *err = MP3D_E_PARAM; scan_job_free(J); return nullptr; }

    mp3dec_batch_plan *P = new mp3dec_batch_plan();
    P->ctx = ctx;
    P->opts = opts;
    P->stream = opts.cuda_stream ? (cudaStream_t)opts.cuda_stream : ctx->stream;
    P->sample_size = opts.float_output ? 4 : 2;
    P->bufs = acquire_set(ctx, 1, (size_t)raw_bytes);
    BufferSet &B = *P->bufs;
    J->P = P;
    /* the scans of a call's chunks must not queue up behind each other (each walks its streams frame by frame at
     * memory latency): every scan plan brings its own stream for the upload, the scan and the record kernels */
    if (opts.cuda_stream) P->scan_stream = nullptr;
    else if (cudaStreamCreateWithFlags(&P->scan_stream, cudaStreamNonBlocking) != cudaSuccess) P->scan_stream = nullptr;
    cudaStream_t up = P->scan_stream ? P->scan_stream : P->stream;
    J->up = up;
    J->t_issue = now_ms();
    if (nc)
    {
        bool ok = (direct || B.h_md.ensure(raw_bytes + 64)) && B.d_raw.ensure(raw_bytes + 64) && B.h_scan.ensure((size_t)nc*sizeof(ScanStream)) &&
                  B.d_scan.ensure((size_t)nc*sizeof(ScanStream)) &&
                  B.d_scan_frames.ensure((size_t)n_slots*sizeof(ScanFrame) + 64) && B.h_sum.ensure((size_t)nc*sizeof(ScanSummary));
        if (!ok) { *err = MP3D_E_MEMORY; mp3dec_batch_plan_destroy(P); scan_job_free(J); return nullptr; }
        ScanStream *hs = (ScanStream *)B.h_scan.p;
        uint32_t slot = 0;
        for (int k = 0; k < nc; k++)
        {
            const ScanHead &H = heads[cand[k]];
            hs[k].src_off = raw_base[k];
            hs[k].size = (uint32_t)H.size;
            hs[k].slot_base = slot;
            hs[k].slot_cap = H.slot_cap;
            memcpy(hs[k].first_hdr, H.p, 4);
            slot += H.slot_cap;
        }
        if (direct)
        {
            /* the caller vouches that the stream buffers sit in page-locked memory (opts.pinned_inputs): the DMA
             * engine reads them where they are, the host touches nothing but the stream heads */
            for (int k = 0; k < nc; k++) hs[k].src_off = (uint64_t)(heads[cand[k]].p - slab_lo);
            cudaMemcpyAsync(B.d_raw.p, slab_lo, raw_bytes, cudaMemcpyHostToDevice, up);
        } else
        {
            uint8_t *h_raw = (uint8_t *)B.h_md.p;
            parallel_for(nc, opts.host_threads, [&](int k) {
                const ScanHead &H = heads[cand[k]];
                memcpy(h_raw + raw_base[k], H.p, H.size);
                memset(h_raw + raw_base[k] + H.size, 0, (size_t)(round_up(H.size + 16, 16) - H.size));
            });
            cudaMemcpyAsync(B.d_raw.p, h_raw, raw_bytes, cudaMemcpyHostToDevice, up);
        }
        cudaMemcpyAsync(B.d_scan.p, hs, (size_t)nc*sizeof(ScanStream), cudaMemcpyHostToDevice, up);
        ScanArgs sa;
        sa.raw = (const uint8_t *)B.d_raw.p;
        sa.streams = (const ScanStream *)B.d_scan.p;
        sa.n_streams = (uint32_t)nc;
        sa.frames = (ScanFrame *)B.d_scan_frames.p;
        /* the summaries go straight into page-locked host memory (unified addressing: the kernel stores through the
         * host pointer).  As a device-to-host copy they would queue in the copy engine behind the PCM read-backs of
         * the earlier chunks -- measured: the scan of chunk 5 of 8 came back 13 ms late, and the first PCM copy waited
         * 0.8 ms for summary copies submitted before it */
        sa.summaries = (ScanSummary *)B.h_sum.p;
        cudaError_t e = launch_stream_scan(sa, up);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&J->ev_sum, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventRecord(J->ev_sum, up);
        if (e != cudaSuccess)
        {
            fprintf(stderr, "minimp3_b200: frame scan failed: %s\n", cudaGetErrorString(e));
            *err = MP3D_E_CUDA;
            mp3dec_batch_plan_destroy(P);
            scan_job_free(J);
            return nullptr;
        }
    }
    J->t_issue = now_ms();
    return J;
}

/* second half: summaries in hand -> where every stream's records go, the record kernels, side info, lane order.
 * Streams the scan handed back are appended to *rest (indices into the job's streams). */
static mp3dec_batch_plan_t *scan_finish(ScanJob *J, int *err, std::vector<int> *rest)
{
    *err = 0;
    mp3dec_batch_plan *P = J->P;
    const mp3dec_batch_opts_t opts = J->opts;
    DeviceCtx *ctx = P->ctx;
    cudaSetDevice(ctx->device);
    BufferSet &B = *P->bufs;
    const int n_streams = J->n_streams, nc = J->nc;
    const std::vector<ScanHead> &heads = J->heads;
    const std::vector<int> &cand = J->cand;
    const uint64_t raw_bytes = J->raw_bytes, n_slots = J->n_slots;
    cudaStream_t up = J->up;
    const double t0 = J->t0, t_heads = J->t_heads, t_issue = J->t_issue;
    *rest = J->rest;
    std::vector<ScanSummary> sums(nc);
    const double t_wait0 = now_ms();
    if (nc)
    {
        const cudaError_t e = cudaEventSynchronize(J->ev_sum);
        if (e != cudaSuccess)
        {
            fprintf(stderr, "minimp3_b200: frame scan failed: %s\n", cudaGetErrorString(e));
            *err = MP3D_E_CUDA;
            mp3dec_batch_plan_destroy(P);
            scan_job_free(J);
            return nullptr;
        }
        memcpy(sums.data(), B.h_sum.p, (size_t)nc*sizeof(ScanSummary));
        /* streams that end in a cut frame or a little junk: the reference restarts its frame search on those last
         * bytes; so do we, with the same search.  Nothing found (the rule): the stream simply ends there. */
        parallel_for(nc, opts.host_threads, [&](int k) {
            ScanSummary &S = sums[k];
            if (S.status != SCAN_TAIL) return;
            const ScanHead &H = heads[cand[k]];
            int free_format_bytes = 0, frame_size = 0;
            find_frame(H.p + S.tail_pos, (int)(H.size - S.tail_pos), &free_format_bytes, &frame_size);
            S.status = frame_size ? SCAN_IRREGULAR : SCAN_OK;
        });
    }
    const double t_scan = now_ms();

    /* the plan's streams: what the scan took + the streams with nothing to decode (they only get zeroed results) */
    std::vector<int> cand_of;          /* plan stream -> candidate index or -1 */
    for (int i = 0, k = 0; i < n_streams; i++)
    {
        const bool is_cand = heads[i].kind == 1;
        if (is_cand && sums[k].status != SCAN_OK) rest->push_back(i);
        else if (heads[i].kind != 2)
        {
            P->out_index.push_back(i);
            cand_of.push_back(is_cand ? k : -1);
        }
        if (is_cand) k++;
    }
    std::sort(rest->begin(), rest->end());
    const int np = (int)P->out_index.size();
    P->n_streams = np;
    P->plans.assign(np, StreamPlan());
    P->pcm_base.assign(np + 1, 0);
    P->gc_base.assign(np + 1, 0);
    if (!B.h_bases.ensure((size_t)std::max(nc, 1)*sizeof(StreamBase)) || !B.d_bases.ensure((size_t)std::max(nc, 1)*sizeof(StreamBase)))
    { *err = MP3D_E_MEMORY; mp3dec_batch_plan_destroy(P); scan_job_free(J); return nullptr; }
    StreamBase *hb = (StreamBase *)B.h_bases.p;
    for (int k = 0; k < nc; k++) { memset(&hb[k], 0, sizeof(StreamBase)); }
    uint64_t gcs = 0, frs = 0, tiles = 0, samples = 0, ists = 0, orders = 0, md = 0, pads = 0;
    uint32_t max_slots = 0;
    for (int j = 0; j < np; j++)
    {
        P->pcm_base[j] = samples;
        P->gc_base[j] = gcs;
        const int k = cand_of[j];
        if (k < 0) continue;
        const ScanHead &H = heads[P->out_index[j]];
        const ScanSummary &S = sums[k];
        StreamPlan &sp = P->plans[j];
        const uint64_t per_frame = (uint64_t)hd_frame_samples(H.p)*S.nch;
        sp.samples_decoded = per_frame*S.n_dec;
        sp.frames = S.n_frames;
        sp.frames_decoded = S.n_dec;
        sp.main_data_bytes = S.md_len;
        sp.avg_bitrate_kbps = S.n_dec ? (int)(S.kbps_sum/S.n_dec) : 0;
        if (!opts.raw_frames)
        {
            sp.channels = H.h.first.channels; sp.hz = H.h.first.hz; sp.layer = H.h.first.layer;
            sp.out_skip = std::min<uint64_t>(H.h.to_skip, sp.samples_decoded);
            sp.out_samples = sp.samples_decoded - sp.out_skip;
            if (H.h.detected_samples && sp.out_samples > H.h.detected_samples) sp.out_samples = H.h.detected_samples;
        } else
        {
            if (S.n_dec) { sp.channels = S.nch; sp.hz = (int)hd_sample_rate_hz(H.p); sp.layer = 3; }
            sp.out_samples = sp.samples_decoded;
        }
        const uint32_t n_gc = S.n_dec*S.n_gr*S.nch, tg = S.nch == 1 ? MP3B_TILE_GRANULES_MONO : MP3B_TILE_GRANULES;
        StreamBase &b = hb[k];
        b.take = 1; b.nch = S.nch; b.n_gr = S.n_gr;
        b.n_frames = S.n_frames; b.n_dec = S.n_dec;
        b.md_base = md; b.pcm_base = samples;
        b.gc_base = (uint32_t)gcs; b.gc_pad_before = (uint32_t)pads; b.fr_base = (uint32_t)frs;
        b.tile_base = (uint32_t)tiles; b.ist_base = (uint32_t)ists;
        max_slots = std::max<uint32_t>(max_slots, (uint32_t)round_up(n_gc, 2));
        pads += round_up(n_gc, 2) - n_gc;
        gcs += round_up(n_gc, 2);
        frs += S.n_dec;
        tiles += ((uint64_t)S.n_dec*S.n_gr + tg - 1)/tg;
        samples += sp.samples_decoded;
        ists += S.n_ist;
        orders += n_gc;
        md += round_up((size_t)S.md_len + 16, 16);
        P->n_frames += S.n_dec;
        P->input_bytes += S.md_len;
    }
    /* the kernels find a frame's / a tile's stream by binary search over these bases: streams that were dropped take
     * the base of the next taken one (they own nothing), so that the arrays stay monotonic */
    uint64_t scan_frames = 0;
    for (int k = 0; k < nc; k++)
    {
        hb[k].frame_base = (uint32_t)scan_frames;
        if (hb[k].take) scan_frames += hb[k].n_frames;
    }
    {
        uint32_t tb = (uint32_t)tiles, fb = (uint32_t)scan_frames;
        for (int k = nc - 1; k >= 0; k--)
        {
            if (hb[k].take) { tb = hb[k].tile_base; fb = hb[k].frame_base; }
            else { hb[k].tile_base = tb; hb[k].frame_base = fb; }
        }
    }
    P->pcm_base[np] = samples;
    P->gc_base[np] = gcs;
    P->md_bytes = md + 4096;
    P->n_gc = gcs; P->n_tiles = tiles; P->n_samples = samples; P->n_ist = ists; P->n_order = orders;
    if (gcs >= (1ull << 31)) { *err = MP3D_E_PARAM; mp3dec_batch_plan_destroy(P); scan_job_free(J); return nullptr; }

    const bool sort_order = frs && orders >= kMinOrderSort && np >= 2;
    const size_t order_scratch = sort_order ? order_sort_scratch_bytes((uint32_t)orders, (uint32_t)np) : 0;
    bool ok = B.d_md.ensure(P->md_bytes) && B.d_gcs.ensure((size_t)gcs*sizeof(GcDesc) + 64) && B.d_frames.ensure((size_t)frs*sizeof(FrameRec) + 64) &&
              B.d_tiles.ensure((size_t)tiles*sizeof(TileDesc) + 64) && B.d_tile_recs.ensure((size_t)tiles*sizeof(TileRec) + 256) &&
              B.d_tile_jobs.ensure((size_t)tiles*MP3B_TILE_JOBS*2 + 256) && B.d_counter.ensure(256) && B.d_istlist.ensure((size_t)ists*4 + 64) &&
              B.d_order.ensure((size_t)orders*4 + 64) &&
              (!sort_order || (B.d_order_sorted.ensure((size_t)orders*4 + 64) && B.d_order_scratch.ensure(order_scratch) &&
                               B.h_gcbase.ensure(((size_t)np + 1)*4 + 64) && B.d_gcbase.ensure(((size_t)np + 1)*4 + 64))) &&
              B.d_xr.ensure((size_t)gcs*576*sizeof(float) + 64) && B.d_nzl.ensure((size_t)gcs*2 + 64) &&
              B.d_istpos.ensure((size_t)gcs*40 + 64) && B.d_pcm.ensure((size_t)samples*P->sample_size + 64);
    if (!ok) { *err = MP3D_E_MEMORY; mp3dec_batch_plan_destroy(P); scan_job_free(J); return nullptr; }
    if (nc) cudaMemcpyAsync(B.d_bases.p, hb, (size_t)nc*sizeof(StreamBase), cudaMemcpyHostToDevice, up);
    if (sort_order)
    {
        uint32_t *gb = (uint32_t *)B.h_gcbase.p;
        for (int j = 0; j <= np; j++) gb[j] = (uint32_t)P->gc_base[j];
        cudaMemcpyAsync(B.d_gcbase.p, gb, ((size_t)np + 1)*4, cudaMemcpyHostToDevice, up);
    }
    cudaMemsetAsync(B.d_md.p, 0, P->md_bytes, up);            /* slack between the streams and the guard behind them */
    if (gcs) cudaMemsetAsync(B.d_gcs.p, 0, (size_t)gcs*sizeof(GcDesc), up);   /* padding slots are never written */
    if (nc && frs)
    {
        EmitArgs ea;
        ea.raw = (const uint8_t *)B.d_raw.p;
        ea.streams = (const ScanStream *)B.d_scan.p;
        ea.bases = (const StreamBase *)B.d_bases.p;
        ea.n_streams = (uint32_t)nc;
        ea.frames = (const ScanFrame *)B.d_scan_frames.p;
        ea.n_frames = (uint32_t)scan_frames;
        ea.frame_recs = (FrameRec *)B.d_frames.p;
        ea.order = (uint32_t *)B.d_order.p;
        ea.ist_list = (uint32_t *)B.d_istlist.p;
        ea.md = (uint8_t *)B.d_md.p;
        ea.tiles = (TileDesc *)B.d_tiles.p;
        ea.n_tiles = (uint32_t)tiles;
        if (launch_stream_emit(ea, up) != cudaSuccess) { *err = MP3D_E_CUDA; mp3dec_batch_plan_destroy(P); scan_job_free(J); return nullptr; }
    }
    if (up != P->stream)
    {
        cudaEventCreateWithFlags(&P->ev_h2d, cudaEventDisableTiming);
        cudaEventRecord(P->ev_h2d, up);
        cudaStreamWaitEvent(P->stream, P->ev_h2d, 0);
    }
    P->sort_order = sort_order; P->order_scratch = order_scratch; P->max_slots = max_slots; P->n_frame_recs = frs;
    if (plan_prep(P) != 0) { *err = MP3D_E_CUDA; mp3dec_batch_plan_destroy(P); scan_job_free(J); return nullptr; }
    P->stats.input_bytes = P->input_bytes;
    P->stats.samples = samples;
    P->stats.n_granule_channels = gcs;
    P->stats.n_tiles = tiles;
    P->stats.n_istereo_granules = ists;
    P->stats.n_frames = P->n_frames;
    P->stats.h2d_bytes = raw_bytes + (uint64_t)nc*(sizeof(ScanStream) + sizeof(StreamBase));
    P->stats.kernels_per_run = (orders ? 1 : 0) + (ists ? 1 : 0) + (tiles ? 2 : 0);
    P->stats.kernels_prep = (nc ? 4 : 0) + (frs ? 1 : 0) + (sort_order ? 2 : 0);
    P->stats.host_parse_ms = (t_issue - t0) + (now_ms() - t_scan);     /* host work; the wait for the scan is not */
    P->stats.h2d_ms = 0;
    static const bool dbg = getenv("MP3B_DEBUG_TIMING") != nullptr;
    if (dbg) fprintf(stderr, "  scan plan: %d streams (%d candidates, %d taken+empty), heads %.3f ms, staging+issue %.3f ms, scan wait %.3f ms, records %.3f ms\n",
                     n_streams, nc, np, t_heads - t0, t_issue - t_heads, t_scan - t_wait0, now_ms() - t_scan);
    scan_job_free(J);
    return P;
}

/* both halves in a row (plans, single chunks) */
static mp3dec_batch_plan_t *plan_create_scan(const mp3dec_batch_stream_t *streams, int n_streams, const mp3dec_batch_opts_t *opts_in,
                                             int *err, std::vector<int> *rest)
{
    ScanJob *J = scan_begin(streams, n_streams, opts_in, err);
    return J ? scan_finish(J, err, rest) : nullptr;
}

int mp3dec_batch_plan_enable_taps(mp3dec_batch_plan_t *P)
{
    if (!P) return MP3D_E_PARAM;
    cudaSetDevice(P->ctx->device);
    BufferSet &B = *P->bufs;
    const size_t n = (size_t)P->n_gc;
    if (!B.d_tap_q.ensure(n*576*2 + 64) || !B.d_tap_scf.ensure(n*40*4 + 64) || !B.d_tap_sub.ensure(n*576*4 + 64) ||
        !B.d_phase_clk.ensure((size_t)P->n_tiles*128 + 64)) return MP3D_E_MEMORY;
    cudaMemsetAsync(B.d_phase_clk.p, 0, (size_t)P->n_tiles*128, P->stream);
    P->taps = true;
    return 0;
}

static int plan_run_stages(mp3dec_batch_plan_t *P, int stage_mask);

int mp3dec_batch_plan_run(mp3dec_batch_plan_t *P) { return plan_run_stages(P, 15); }

/* measurement aid: launch a single stage (0 entropy, 1 intensity stereo, 2 transform, 3 Layer I/II, 4 the plan's GPU
 * preparation: side-info parse + lane order) on the plan's stream */
int mp3dec_batch_plan_run_stage(mp3dec_batch_plan_t *P, int stage)
{
    if (stage < 0 || stage > 4) return MP3D_E_PARAM;
    if (stage == 4)
    {
        if (!P) return MP3D_E_PARAM;
        cudaSetDevice(P->ctx->device);
        return plan_prep(P);
    }
    return plan_run_stages(P, 1 << stage);
}

static int plan_run_stages(mp3dec_batch_plan_t *P, int stage_mask)
{
    if (!P) return MP3D_E_PARAM;
    cudaSetDevice(P->ctx->device);
    BufferSet &B = *P->bufs;
    if (P->taps)
    {
        cudaMemsetAsync(B.d_tap_q.p, 0, (size_t)P->n_gc*576*2, P->stream);
        cudaMemsetAsync(B.d_tap_scf.p, 0, (size_t)P->n_gc*40*4, P->stream);
        cudaMemsetAsync(B.d_tap_sub.p, 0, (size_t)P->n_gc*576*4, P->stream);
        cudaMemsetAsync(B.d_istpos.p, 0, (size_t)P->n_gc*40, P->stream);
    }
    EntropyArgs ea;
    memset(&ea, 0, sizeof(ea));
    ea.md_words = (const uint32_t *)B.d_md.p;
    ea.md_n_words = P->md_bytes/4;
    ea.gcs = (const GcDesc *)B.d_gcs.p;
    ea.n_gc = (uint32_t)P->n_gc;
    ea.order = (const uint32_t *)(P->order_sorted ? B.d_order_sorted.p : B.d_order.p);
    ea.n_order = (uint32_t)P->n_order;
    ea.group_counter = (uint32_t *)B.d_counter.p;
    ea.xr = (float *)B.d_xr.p;
    ea.nzl = (uint16_t *)B.d_nzl.p;
    ea.ist_pos = (uint8_t *)B.d_istpos.p;
    static const bool clk_only = getenv("MP3B_DEBUG_CLK_ONLY") != nullptr;   /* tools/phase_times.py: phase clocks without the stage taps */
    const bool stage_taps = P->taps && !clk_only;
    ea.tap_q = stage_taps ? (int16_t *)B.d_tap_q.p : nullptr;
    ea.tap_scf = stage_taps ? (float *)B.d_tap_scf.p : nullptr;
    cudaError_t e = cudaSuccess;
    if (stage_mask & 1) e = launch_entropy(P->ctx->tables, ea, P->stream);
    if (e == cudaSuccess && P->n_ist && (stage_mask & 2))
    {
        StereoArgs sa;
        sa.gcs = ea.gcs;
        sa.granules = (const uint32_t *)B.d_istlist.p;
        sa.n_granules = (uint32_t)P->n_ist;
        sa.xr = ea.xr;
        sa.nzl = ea.nzl;
        sa.ist_pos = ea.ist_pos;
        e = launch_intensity_stereo(P->ctx->tables, sa, P->stream);
    }
    if (e == cudaSuccess && (stage_mask & 4))
    {
        TransformArgs ta;
        memset(&ta, 0, sizeof(ta));
        ta.gcs = ea.gcs;
        ta.tiles = (const TileDesc *)B.d_tiles.p;
        ta.n_tiles = (uint32_t)P->n_tiles;
        ta.xr = ea.xr;
        ta.nzl = ea.nzl;
        ta.tile_recs = (TileRec *)B.d_tile_recs.p;
        ta.tile_jobs = (uint16_t *)B.d_tile_jobs.p;
        ta.pcm = B.d_pcm.p;
        ta.float_output = P->opts.float_output;
        ta.tap_sub = stage_taps ? (float *)B.d_tap_sub.p : nullptr;
        ta.phase_clk = P->taps ? (long long *)B.d_phase_clk.p : nullptr;
        e = launch_transform(P->ctx->tables, ta, P->stream);
    }
    if (e == cudaSuccess && P->n_l12 && (stage_mask & 8))
    {
        Layer12Args la;
        la.md_words = ea.md_words; la.md_n_words = ea.md_n_words;
        la.frames = (const L12FrameDesc *)B.d_l12.p; la.n_frames = (uint32_t)P->n_l12;
        la.sub12 = (float *)B.d_sub12.p;
        la.info = (L12Compact *)B.d_l12info.p;
        e = launch_l12_dequant(la, P->stream);
        if (e == cudaSuccess)
        {
            Synth12Args sa;
            memset(&sa, 0, sizeof(sa));
            sa.tiles = (const TileDesc *)B.d_tiles12.p; sa.n_tiles = (uint32_t)P->n_tiles12;
            sa.sub12 = la.sub12; sa.pcm = B.d_pcm.p; sa.float_output = P->opts.float_output;
            e = launch_l12_synth(P->ctx->tables, sa, P->stream);
        }
    }
    if (e != cudaSuccess)
    {
        fprintf(stderr, "minimp3_b200: kernel launch failed: %s\n", cudaGetErrorString(e));
        return MP3D_E_CUDA;
    }
    return 0;
}

int mp3dec_batch_plan_sync(mp3dec_batch_plan_t *P)
{
    if (!P) return MP3D_E_PARAM;
    cudaSetDevice(P->ctx->device);
    cudaError_t e = cudaStreamSynchronize(P->stream);
    if (e != cudaSuccess)
    {
        fprintf(stderr, "minimp3_b200: stream failed: %s\n", cudaGetErrorString(e));
        return MP3D_E_CUDA;
    }
    return 0;
}

void *mp3dec_batch_plan_device_pcm(mp3dec_batch_plan_t *P) { return P ? P->bufs->d_pcm.p : nullptr; }

int mp3dec_batch_plan_read_tap(mp3dec_batch_plan_t *P, int which, void *dst, size_t dst_bytes)
{
    if (!P || !dst) return MP3D_E_PARAM;
    cudaSetDevice(P->ctx->device);
    BufferSet &B = *P->bufs;
    const size_t n = (size_t)P->n_gc;
    const void *src = nullptr;
    size_t bytes = 0;
    switch (which)
    {
    case 0: src = B.d_tap_q.p; bytes = n*576*2; break;
    case 1: src = B.d_tap_scf.p; bytes = n*40*4; break;
    case 2: src = B.d_xr.p; bytes = n*576*4; break;
    case 3: src = B.d_tap_sub.p; bytes = n*576*4; break;
    case 4: src = B.d_gcs.p; bytes = n*sizeof(GcDesc); break;
    case 5: src = B.d_istpos.p; bytes = n*40; break;
    case 6: src = B.d_nzl.p; bytes = n*2; break;
    case 7: src = B.d_phase_clk.p; bytes = (size_t)P->n_tiles*128; break;
    case 8: src = P->order_sorted ? B.d_order_sorted.p : B.d_order.p; bytes = (size_t)P->n_order*4; break;
    default: return MP3D_E_PARAM;
    }
    if ((which == 0 || which == 1 || which == 3 || which == 7) && !P->taps) return MP3D_E_PARAM;
    if (dst_bytes < bytes) return MP3D_E_PARAM;
    if (cudaStreamSynchronize(P->stream) != cudaSuccess) return MP3D_E_CUDA;
    if (bytes && cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return MP3D_E_CUDA;
    return 0;
}

/* starts the PCM read-back behind the kernels already queued on the plan's stream */
static int plan_begin_fetch(mp3dec_batch_plan_t *P, cudaStream_t copy_stream)
{
    BufferSet &B = *P->bufs;
    const size_t ss = P->sample_size;
    if (P->opts.output == MP3D_BATCH_OUT_DEVICE || !P->n_samples) { P->d2h_started = true; return 0; }
    if (!B.h_pcm.ensure((size_t)P->n_samples*ss + 64)) return MP3D_E_MEMORY;
    if (copy_stream != P->stream)
    {
        if (!P->ev_kernels) cudaEventCreateWithFlags(&P->ev_kernels, getenv("MP3B_DEBUG_TIMING") ? cudaEventDefault : cudaEventDisableTiming);
        cudaEventRecord(P->ev_kernels, P->stream);
        cudaStreamWaitEvent(copy_stream, P->ev_kernels, 0);
    }
    static const bool dbg = getenv("MP3B_DEBUG_TIMING") != nullptr;
    if (dbg)
    {   /* debug timeline: when this plan's read-back ran, relative to the end of its kernels */
        cudaEventCreate(&P->ev_dbg0);
        cudaEventRecord(P->ev_dbg0, copy_stream);
    }
    cudaMemcpyAsync(B.h_pcm.p, B.d_pcm.p, (size_t)P->n_samples*ss, cudaMemcpyDeviceToHost, copy_stream);
    if (!P->ev_d2h) cudaEventCreateWithFlags(&P->ev_d2h, dbg ? cudaEventDefault : cudaEventDisableTiming);
    cudaEventRecord(P->ev_d2h, copy_stream);
    P->d2h_started = true;
    return 0;
}

static int plan_finish_fetch(mp3dec_batch_plan_t *P, mp3dec_file_info_t *infos, int *rets)
{
    BufferSet &B = *P->bufs;
    const size_t ss = P->sample_size;
    const int mode = P->opts.output;
    uint8_t *host = (uint8_t *)B.h_pcm.p;
    cudaError_t e = P->ev_d2h ? cudaEventSynchronize(P->ev_d2h) : cudaStreamSynchronize(P->stream);
    if (e != cudaSuccess)
    {
        fprintf(stderr, "minimp3_b200: decode failed: %s\n", cudaGetErrorString(e));
        return MP3D_E_CUDA;
    }
    if (P->ev_dbg0)
    {
        float ms = 0.f, k1 = -1.f, c0 = -1.f, c1 = -1.f;
        cudaEventElapsedTime(&ms, P->ev_dbg0, P->ev_d2h);
        if (g_dbg_base)
        {
            if (P->ev_kernels) cudaEventElapsedTime(&k1, g_dbg_base, P->ev_kernels);
            cudaEventElapsedTime(&c0, g_dbg_base, P->ev_dbg0);
            cudaEventElapsedTime(&c1, g_dbg_base, P->ev_d2h);
        }
        fprintf(stderr, "  read-back of %zu bytes: %.3f ms = %.1f GB/s; kernels done %.3f, copy queue free %.3f, copy done %.3f\n",
                (size_t)P->n_samples*ss, ms, (double)P->n_samples*ss/ms/1e6, k1, c0, c1);
        cudaEventDestroy(P->ev_dbg0);
        P->ev_dbg0 = nullptr;
    }
    std::atomic<int> oom(0);
    parallel_for(P->n_streams, mode == MP3D_BATCH_OUT_MALLOC ? P->opts.host_threads : 1, [&](int i) {
        const StreamPlan &sp = P->plans[i];
        const int o = P->out_index.empty() ? i : P->out_index[i];
        if (rets) rets[o] = sp.ret;
        if (!infos) return;
        mp3dec_file_info_t &fi = infos[o];
        memset(&fi, 0, sizeof(fi));
        fi.samples = (size_t)sp.out_samples;
        fi.channels = sp.channels; fi.hz = sp.hz; fi.layer = sp.layer; fi.avg_bitrate_kbps = sp.avg_bitrate_kbps;
        const size_t first = (size_t)(P->pcm_base[i] + sp.out_skip)*ss;
        if (!sp.out_samples) return;
        if (mode == MP3D_BATCH_OUT_DEVICE) fi.buffer = (mp3d_sample_t *)((uint8_t *)B.d_pcm.p + first);
        else if (mode == MP3D_BATCH_OUT_PINNED) fi.buffer = (mp3d_sample_t *)(host + first);
        else
        {
            void *m = malloc((size_t)sp.out_samples*ss);
            if (!m) { oom = 1; fi.samples = 0; if (rets) rets[o] = MP3D_E_MEMORY; return; }
            memcpy(m, host + first, (size_t)sp.out_samples*ss);
            fi.buffer = (mp3d_sample_t *)m;
        }
    });
    return oom ? MP3D_E_MEMORY : 0;
}

int mp3dec_batch_plan_fetch(mp3dec_batch_plan_t *P, mp3dec_file_info_t *infos, int *rets)
{
    if (!P) return MP3D_E_PARAM;
    cudaSetDevice(P->ctx->device);
    if (!P->d2h_started)
    {
        const int r = plan_begin_fetch(P, P->stream);
        if (r) return r;
    }
    const int r = plan_finish_fetch(P, infos, rets);
    P->d2h_started = false;
    return r;
}

void mp3dec_batch_plan_stats(const mp3dec_batch_plan_t *P, mp3dec_batch_stats_t *out)
{
    if (P && out) *out = P->stats;
}

void mp3dec_batch_plan_destroy(mp3dec_batch_plan_t *P)
{
    if (!P) return;
    if (P->ev_kernels) cudaEventDestroy(P->ev_kernels);
    if (P->ev_d2h) cudaEventDestroy(P->ev_d2h);
    if (P->ev_h2d) cudaEventDestroy(P->ev_h2d);
    if (P->scan_stream) cudaStreamDestroy(P->scan_stream);
    if (P->bufs)
    {
        if (P->export_on_destroy)
        {
            std::lock_guard<std::mutex> lock(P->ctx->mu);
            P->ctx->exported_sets.push_back(P->bufs);
        } else
            release_set(P->ctx, P->bufs);
    }
    delete P;
}

/* Whole batch in one call.  Large batches are cut into chunks of streams and pipelined: while chunk c is being
 * decoded and its PCM travels back over PCIe (second stream), the host threads already walk chunk c+1 and its
 * bitstream goes up.  Per-stream results are independent of the chunking. */
static thread_local mp3dec_batch_call_stats_t t_call_stats;

void mp3dec_batch_last_call_stats(mp3dec_batch_call_stats_t *out)
{
    if (out) *out = t_call_stats;
}

void mp3dec_batch_release_outputs(int device)
{
    DeviceCtx *c = get_ctx(device);
    if (c) release_exported(c);
}

/* opts->n_devices > 1: the streams are split over the listed GPUs by greedy LPT on input bytes (mp3dec_batch_shard)
 * and every shard goes through the single-device call on its own host thread with its share of the host cores;
 * results land in the caller's arrays at the streams' own indices.  No data-path collective: streams are
 * independent. */
static int decode_batch_multi(const mp3dec_batch_stream_t *streams, int n_streams, mp3dec_file_info_t *infos_out,
                              int *rets_out, const mp3dec_batch_opts_t *opts)
{
    const int nd = std::min(opts->n_devices, 8);
    const double call_t0 = now_ms();
    for (int d = 0; d < nd; d++)
        if (!get_ctx(opts->devices[d])) return MP3D_E_CUDA;
    std::vector<size_t> sizes(n_streams);
    for (int i = 0; i < n_streams; i++) sizes[i] = streams[i].size;
    std::vector<int> shard_of(n_streams);
    if (opts->pinned_inputs && opts->pinned_inputs_bytes)
    {
        /* streams in a caller-pinned slab are uploaded where they are, as one span per chunk: give every device a
         * contiguous run of streams (equal bytes) instead of the LPT scatter */
        uint64_t total = 0, acc = 0;
        for (int i = 0; i < n_streams; i++) total += sizes[i];
        for (int i = 0; i < n_streams; i++)
        {
            shard_of[i] = (int)std::min<uint64_t>((uint64_t)nd - 1, total ? (acc + sizes[i]/2)*(uint64_t)nd/total : 0);
            acc += sizes[i];
        }
    } else
        mp3dec_batch_shard(sizes.data(), n_streams, nd, shard_of.data());
    std::vector<std::vector<int>> ids(nd);
    for (int i = 0; i < n_streams; i++) ids[shard_of[i]].push_back(i);
    /* one pool per device slot, each with an equal share of the cores; created on first use, kept */
    static std::mutex pools_mu;
    static std::vector<HostPool *> pools[9];
    int cores = opts->host_threads > 0 ? opts->host_threads : (int)std::thread::hardware_concurrency();
    if (cores < nd) cores = nd;
    {
        std::lock_guard<std::mutex> lk(pools_mu);
        while ((int)pools[nd].size() < nd) pools[nd].push_back(new HostPool(std::max(1, cores/nd)));
    }
    std::vector<int> errs(nd, 0);
    std::vector<mp3dec_batch_call_stats_t> stats(nd);
    std::vector<std::thread> th;
    for (int d = 0; d < nd; d++)
        th.emplace_back([&, d] {
            HostPool::set_current(pools[nd][d]);
            const int n = (int)ids[d].size();
            std::vector<mp3dec_batch_stream_t> sub(n);
            std::vector<mp3dec_file_info_t> infos(infos_out ? n : 0);
            std::vector<int> rets(n, 0);
            for (int k = 0; k < n; k++) sub[k] = streams[ids[d][k]];
            mp3dec_batch_opts_t o = *opts;
            o.n_devices = 0;
            o.device = opts->devices[d];
            o.host_threads = 0;
            errs[d] = n ? mp3dec_decode_batch_cuda(sub.data(), n, infos_out ? infos.data() : nullptr, rets.data(), &o) : 0;
            mp3dec_batch_last_call_stats(&stats[d]);
            if (!n) memset(&stats[d], 0, sizeof(stats[d]));
            for (int k = 0; k < n; k++)
            {
                if (infos_out) infos_out[ids[d][k]] = infos[k];
                if (rets_out) rets_out[ids[d][k]] = rets[k];
            }
            HostPool::set_current(nullptr);
        });
    for (auto &t : th) t.join();
    mp3dec_batch_call_stats_t &cs = t_call_stats;
    memset(&cs, 0, sizeof(cs));
    int err = 0;
    for (int d = 0; d < nd; d++)
    {
        if (errs[d] && !err) err = errs[d];
        cs.host_parse_ms = std::max(cs.host_parse_ms, stats[d].host_parse_ms);
        cs.issue_ms = std::max(cs.issue_ms, stats[d].issue_ms);
        cs.finish_ms = std::max(cs.finish_ms, stats[d].finish_ms);
        cs.h2d_bytes += stats[d].h2d_bytes; cs.d2h_bytes += stats[d].d2h_bytes;
        cs.n_chunks += stats[d].n_chunks;
        cs.scanned_streams += stats[d].scanned_streams;
    }
    cs.n_devices = nd;
    cs.total_ms = now_ms() - call_t0;
    return err;
}

int mp3dec_decode_batch_cuda(const mp3dec_batch_stream_t *streams, int n_streams, mp3dec_file_info_t *infos_out,
                             int *rets_out, const mp3dec_batch_opts_t *opts)
{
    if (!streams || n_streams < 0) return MP3D_E_PARAM;
    if (opts && opts->n_devices > 1 && !opts->cuda_stream) return decode_batch_multi(streams, n_streams, infos_out, rets_out, opts);
    int err = 0;
    const double call_t0 = now_ms();
    mp3dec_batch_call_stats_t &cs = t_call_stats;
    memset(&cs, 0, sizeof(cs));
    const bool user_stream = opts && opts->cuda_stream;
    const bool exports = opts && opts->output != MP3D_BATCH_OUT_MALLOC;
    if (exports)
    {
        /* the pointers of the previous pinned / device-output call on this device end here */
        DeviceCtx *c = get_ctx(opts->device);
        if (c) release_exported(c);
    }
    int n_chunks = 1;
    if (!user_stream && n_streams >= 64)
    {
        size_t total = 0;
        for (int i = 0; i < n_streams; i++) total += streams[i].size;
        n_chunks = (int)std::min<size_t>(8, std::max<size_t>(1, total/(8u << 20)));   /* >= 8 MB of input per chunk */
        n_chunks = std::min(n_chunks, n_streams/32);
        if (n_chunks < 1) n_chunks = 1;
    }
    /* frame sync on the GPU (k_scan.cu): per-frame host work disappears; streams the scan hands back (Layer I/II, free
     * format, anything irregular) go through the host walker in a second plan of the same chunk.
     * opts.gpu_frame_scan: 0 auto (see below), 1 always, -1 never. */
    static const int scan_env = [] { const char *e = getenv("MP3B_GPU_SCAN"); return e ? atoi(e) : 0; }();
    const int scan_mode = scan_env ? scan_env : (opts ? opts->gpu_frame_scan : 0);
    struct Part { mp3dec_batch_plan_t *plan; int first; };
    std::vector<Part> parts;
    std::vector<int> first(n_chunks + 1);
    /* small first chunks get the PCIe read-back going early; the later ones amortise launch overheads */
    {
        double w = 1.0, total_w = 0.0;
        std::vector<double> weight(n_chunks);
        for (int c = 0; c < n_chunks; c++) { weight[c] = w; total_w += w; if (c < 3) w *= 2.0; }
        double acc = 0.0;
        first[0] = 0;
        for (int c = 0; c < n_chunks; c++)
        {
            acc += weight[c];
            first[c + 1] = c + 1 == n_chunks ? n_streams : std::max(first[c] + 1, (int)(n_streams*acc/total_w));
        }
    }
    static const bool dbg_timing = getenv("MP3B_DEBUG_TIMING") != nullptr;
    const double tt0 = now_ms();
    if (dbg_timing)
    {
        DeviceCtx *c = get_ctx(opts ? opts->device : 0);
        if (c)
        {
            if (!g_dbg_base) cudaEventCreate(&g_dbg_base);
            cudaEventRecord(g_dbg_base, c->stream_d2h);
        }
    }
    auto launch = [&](mp3dec_batch_plan_t *P, int first_stream) {
        parts.push_back(Part{ P, first_stream });
        P->export_on_destroy = exports;
        cs.host_parse_ms += P->stats.host_parse_ms;
        cs.h2d_bytes += P->stats.h2d_bytes;
        if (P->opts.output != MP3D_BATCH_OUT_DEVICE) cs.d2h_bytes += P->n_samples*P->sample_size;
        int e = mp3dec_batch_plan_run(P);
        if (!e) e = plan_begin_fetch(P, P->ctx->stream_d2h);
        return e;
    };
    /* all scans of the call are started first: each walks its streams frame by frame at memory latency (~1 ms), and
     * they overlap each other, the uploads of the later chunks and the decode kernels of the earlier ones */
    std::vector<ScanJob *> jobs(n_chunks, nullptr);
    for (int c = 0; c < n_chunks && !err; c++)
    {
        const int n_c = first[c + 1] - first[c];
        /* auto: only where it pays.  With pageable inputs the host has to copy every byte into the staging slab either
         * way -- that copy, not the frame walk, is what the host stage costs (measured: 4.2 ms of 4.5 ms on config 2) --
         * and the scan only adds its latency; with inputs the caller pinned (opts.pinned_inputs) the copy disappears
         * and the host is left with the stream heads. */
        const bool pays = opts && opts->pinned_inputs && opts->pinned_inputs_bytes && n_streams >= 64;   /* the call's count: the first chunks are small by design */
        if (scan_mode > 0 || (scan_mode == 0 && pays)) jobs[c] = scan_begin(streams + first[c], n_c, opts, &err);
    }
    for (int c = 0; c < n_chunks && !err; c++)
    {
        const double a0 = now_ms();
        const int n_c = first[c + 1] - first[c];
        if (jobs[c])
        {
            std::vector<int> rest;
            mp3dec_batch_plan_t *P = scan_finish(jobs[c], &err, &rest);
            jobs[c] = nullptr;
            if (!P) break;
            cs.scanned_streams += n_c - (int)rest.size();
            err = launch(P, first[c]);
            if (!err && !rest.empty())
            {
                std::vector<mp3dec_batch_stream_t> sub(rest.size());
                for (size_t k = 0; k < rest.size(); k++) sub[k] = streams[first[c] + rest[k]];
                mp3dec_batch_plan_t *Q = plan_create_impl(sub.data(), (int)sub.size(), opts, &err, true);
                if (!Q) break;
                Q->out_index = rest;
                err = launch(Q, first[c]);
            }
        } else
        {
            mp3dec_batch_plan_t *P = plan_create_impl(streams + first[c], n_c, opts, &err, true);
            if (!P) break;
            err = launch(P, first[c]);
        }
        if (dbg_timing) fprintf(stderr, "chunk %d: %d streams, issued in %.3f ms, t=%.3f\n", c, n_c, now_ms() - a0, now_ms() - tt0);
    }
    for (ScanJob *J : jobs)
        if (J)
        {   /* an error ended the call early: let the queued work drain, drop the rest */
            cudaStreamSynchronize(J->up);
            mp3dec_batch_plan_destroy(J->P);
            scan_job_free(J);
        }
    const double tt1 = now_ms();
    for (Part &pt : parts)
    {
        if (!err)
        {
            err = plan_finish_fetch(pt.plan, infos_out ? infos_out + pt.first : nullptr, rets_out ? rets_out + pt.first : nullptr);
            if (dbg_timing) fprintf(stderr, "  part done at t=%.3f\n", now_ms() - tt0);
        } else
            cudaStreamSynchronize(pt.plan->stream);
    }
    const double tt2 = now_ms();
    for (Part &pt : parts)
    {
        if (err) cudaStreamSynchronize(pt.plan->ctx->stream_d2h);
        mp3dec_batch_plan_destroy(pt.plan);
    }
    if (dbg_timing) fprintf(stderr, "issue %.3f ms, finish %.3f ms, destroy %.3f ms\n", tt1 - tt0, tt2 - tt1, now_ms() - tt2);
    cs.n_chunks = n_chunks; cs.n_devices = 1;
    cs.issue_ms = tt1 - call_t0; cs.finish_ms = tt2 - tt1; cs.total_ms = now_ms() - call_t0;
    return err;
}

/* ---- drop-in entry points ------------------------------------------------------------------------------- */

void mp3dec_init(mp3dec_t *dec) { dec->header[0] = 0; }   /* minimp3.h:1708-1711 */

int mp3dec_detect_buf(const uint8_t *buf, size_t buf_size) { return detect_buf(buf, buf_size); }

extern "C" __attribute__((visibility("hidden"))) int mp3dec_load_buf_progress(mp3dec_t *dec, const uint8_t *buf, size_t buf_size, mp3dec_file_info_t *info,
                                        MP3D_PROGRESS_CB progress_cb, void *user_data, int float_output);

static int load_buf_impl(mp3dec_t *dec, const uint8_t *buf, size_t buf_size, mp3dec_file_info_t *info,
                         MP3D_PROGRESS_CB progress_cb, void *user_data, int float_output)
{
    if (!dec || !buf || !info || (size_t)-1 == buf_size) return MP3D_E_PARAM;
    /* a progress callback wants the frames one by one and may cancel: frame-by-frame flavour (csrc/ex_api.cpp) */
    if (progress_cb) return mp3dec_load_buf_progress(dec, buf, buf_size, info, progress_cb, user_data, float_output);
    memset(info, 0, sizeof(*info));
    mp3dec_init(dec);
    mp3dec_batch_stream_t s = { buf, buf_size };
    mp3dec_batch_opts_t o;
    memset(&o, 0, sizeof(o));
    o.float_output = float_output;
    o.output = MP3D_BATCH_OUT_MALLOC;
    o.host_threads = 1;
    int ret = 0;
    int err = mp3dec_decode_batch_cuda(&s, 1, info, &ret, &o);
    if (err) return err;
    return ret;
}

int mp3dec_load_buf(mp3dec_t *dec, const uint8_t *buf, size_t buf_size, mp3dec_file_info_t *info,
                    MP3D_PROGRESS_CB progress_cb, void *user_data)
{
    return load_buf_impl(dec, buf, buf_size, info, progress_cb, user_data, 0);
}

int mp3dec_load_buf_f32(mp3dec_t *dec, const uint8_t *buf, size_t buf_size, mp3dec_file_info_t *info,
                        MP3D_PROGRESS_CB progress_cb, void *user_data)
{
    return load_buf_impl(dec, buf, buf_size, info, progress_cb, user_data, 1);
}

/* One frame through the same kernels, with the reference's carried state (mdct_overlap, qmf_state,
 * reservoir bytes) living in the caller's mp3dec_t exactly where the reference keeps it. */
static int decode_frame_impl(mp3dec_t *dec, const uint8_t *mp3, int mp3_bytes, void *pcm, mp3dec_frame_info_t *info,
                             int float_output)
{
    SyncState st;
    memcpy(st.header, dec->header, 4);
    st.free_format_bytes = dec->free_format_bytes;
    /* the reference keeps 0 <= reserv <= 511; anything else is a caller's uninitialised or foreign struct */
    const int reserv_before = std::min(std::max(dec->reserv, 0), 511);
    st.reserv = reserv_before;
    FrameParse fp;
    fp.info = FrameInfo{ info->frame_bytes, info->frame_offset, info->channels, info->hz, info->layer, info->bitrate_kbps };
    parse_frame(&st, mp3, mp3_bytes, &fp, pcm == nullptr);
    if (fp.reset) memset(dec, 0, sizeof(*dec));
    memcpy(dec->header, st.header, 4);
    dec->free_format_bytes = st.free_format_bytes;
    info->frame_bytes = fp.info.frame_bytes; info->frame_offset = fp.info.frame_offset;
    info->channels = fp.info.channels; info->hz = fp.info.hz; info->layer = fp.info.layer;
    info->bitrate_kbps = fp.info.bitrate_kbps;
    if (!fp.hdr) return 0;                              /* no frame found */
    if (!pcm) return (int)hdr_frame_samples(fp.hdr);    /* parse only (minimp3.h:1748-1751) */
    if (fp.init_only) return 0;
    if (fp.layer != 3)
    {
        /* Layer I/II (minimp3.h:1783-1802): dequantise on the GPU, 12-slot synthesis with the carried history */
        DeviceCtx *ctx = get_ctx(0);
        if (!ctx) return 0;
        std::lock_guard<std::mutex> lock(ctx->mu);
        cudaSetDevice(ctx->device);
        BufferSet &B = ctx->frame_set;
        const size_t ss = float_output ? 4 : 2;
        const int n_blocks = fp.layer == 1 ? 1 : 3, nch = fp.nch;
        const int n_samples = n_blocks*384*nch;
        const size_t md_bytes = round_up((size_t)fp.payload_bytes + 64, 16);
        bool ok = B.h_md.ensure(md_bytes + 256) && B.d_md.ensure(md_bytes) && B.d_l12.ensure(64) && B.d_tiles12.ensure(64) &&
                  B.d_sub12.ensure(6*384*4) && B.d_l12info.ensure(sizeof(L12Compact)) && B.d_pcm.ensure(2304*4) && B.h_pcm.ensure(2304*4) &&
                  ctx->d_state_in.ensure(1536*4) && ctx->d_state_out.ensure(1536*4) && ctx->h_state.ensure(2*1536*4);
        if (!ok) return 0;
        uint8_t *h = (uint8_t *)B.h_md.p;
        memset(h, 0, md_bytes);
        memcpy(h, fp.payload, (size_t)fp.payload_bytes);
        L12FrameDesc fd;
        memset(&fd, 0, sizeof(fd));
        memcpy(fd.hdr, fp.hdr, 4);
        fd.kbps = (uint16_t)fp.info.bitrate_kbps; fd.nch = (uint8_t)nch; fd.n_blocks = (uint8_t)n_blocks;
        TileDesc tile;
        memset(&tile, 0, sizeof(tile));
        tile.n_gr = (uint8_t)n_blocks; tile.nch = (uint8_t)nch;
        tile.halo[0][0] = tile.halo[0][1] = tile.halo[1][0] = tile.halo[1][1] = -1;
        memcpy(h + md_bytes, &fd, sizeof(fd));
        memcpy(h + md_bytes + 64, &tile, sizeof(tile));
        float *hs = (float *)ctx->h_state.p;
        memcpy(hs, dec->qmf_state, 960*4);
        cudaStream_t s = ctx->stream;
        cudaMemcpyAsync(B.d_md.p, h, md_bytes, cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(B.d_l12.p, h + md_bytes, sizeof(fd), cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(B.d_tiles12.p, h + md_bytes + 64, sizeof(tile), cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(ctx->d_state_in.p, hs, 960*4, cudaMemcpyHostToDevice, s);
        Layer12Args la;
        la.md_words = (const uint32_t *)B.d_md.p; la.md_n_words = md_bytes/4;
        la.frames = (const L12FrameDesc *)B.d_l12.p; la.n_frames = 1; la.sub12 = (float *)B.d_sub12.p; la.info = (L12Compact *)B.d_l12info.p;
        cudaError_t e = launch_l12_dequant(la, s);
        if (e == cudaSuccess)
        {
            Synth12Args sa;
            memset(&sa, 0, sizeof(sa));
            sa.tiles = (const TileDesc *)B.d_tiles12.p; sa.n_tiles = 1; sa.sub12 = la.sub12; sa.pcm = B.d_pcm.p;
            sa.float_output = float_output;
            sa.state_in = (const float *)ctx->d_state_in.p; sa.state_out = (float *)ctx->d_state_out.p;
            e = launch_l12_synth(ctx->tables, sa, s);
        }
        if (e == cudaSuccess)
        {
            cudaMemcpyAsync(B.h_pcm.p, B.d_pcm.p, (size_t)n_samples*ss, cudaMemcpyDeviceToHost, s);
            cudaMemcpyAsync(hs + 1536, ctx->d_state_out.p, 960*4, cudaMemcpyDeviceToHost, s);
            e = cudaStreamSynchronize(s);
        }
        if (e != cudaSuccess)
        {
            fprintf(stderr, "minimp3_b200: frame decode failed: %s\n", cudaGetErrorString(e));
            return 0;
        }
        memcpy(pcm, B.h_pcm.p, (size_t)n_samples*ss);
        for (int ch = 0; ch < nch; ch++) memcpy(dec->qmf_state + ch*480, hs + 1536 + ch*480, 480*4);
        return (int)hdr_frame_samples(fp.hdr);
    }
    /* L3_restore_reservoir (minimp3.h:1228-1236): main data = tail of the reservoir + this payload */
    const int reserv_old = fp.reset ? 0 : reserv_before;
    const int mdb = fp.main_data_begin;
    const int bytes_have = std::min(reserv_old, mdb);
    uint8_t maindata[511 + 2304 + 64 + 16];      /* copied in 16-byte units below */
    memset(maindata, 0, sizeof(maindata));
    memcpy(maindata, dec->reserv_buf + std::max(0, reserv_old - mdb), (size_t)bytes_have);
    memcpy(maindata + bytes_have, fp.payload, (size_t)fp.payload_bytes);
    const int md_len = bytes_have + fp.payload_bytes;
    int result = 0;

    if (fp.success)
    {
        DeviceCtx *ctx = get_ctx(0);
        if (!ctx) return 0;
        std::lock_guard<std::mutex> lock(ctx->mu);
        cudaSetDevice(ctx->device);
        BufferSet &B = ctx->frame_set;
        const int n_gc = fp.n_gr*fp.nch;
        const size_t md_bytes = round_up((size_t)md_len + 64, 16);
        const size_t ss = float_output ? 4 : 2;
        const int n_samples = fp.n_gr*576*fp.nch;
        /* 4 KB behind the payload: a damaged granule can ask the Huffman decoder for up to 288 pairs of 47 bits */
        const size_t md_cap_before = B.d_md.cap;
        bool ok = B.h_md.ensure(md_bytes + 4*sizeof(GcDesc) + sizeof(TileDesc) + 64) && B.d_md.ensure(md_bytes + 4096) &&
                  B.d_gcs.ensure(4*sizeof(GcDesc)) && B.d_tiles.ensure(sizeof(TileDesc)) && B.d_tile_recs.ensure(sizeof(TileRec)) && B.d_tile_jobs.ensure(MP3B_TILE_JOBS*2) && B.d_counter.ensure(256) && B.d_istlist.ensure(16) && B.d_order.ensure(16) &&
                  B.d_xr.ensure(4*576*4) && B.d_nzl.ensure(16) && B.d_istpos.ensure(4*40) &&
                  B.d_pcm.ensure(2304*4) && B.h_pcm.ensure(2304*4) && ctx->d_state_in.ensure(1536*4) &&
                  ctx->d_state_out.ensure(1536*4) && ctx->h_state.ensure(2*1536*4);
        if (!ok) return 0;
        if (B.d_md.cap != md_cap_before) cudaMemsetAsync(B.d_md.p, 0, B.d_md.cap, ctx->stream);   /* fresh allocation: defined guard bytes (same stream as the copies) */
        uint8_t *h = (uint8_t *)B.h_md.p;
        memcpy(h, maindata, md_bytes);
        GcDesc gcs[4];
        uint32_t ist_list[2];
        int n_ist = 0;
        uint64_t bit = 0;
        for (int igr = 0; igr < fp.n_gr; igr++)
        {
            if ((fp.hdr[3] & 0x10) && fp.nch == 2) ist_list[n_ist++] = (uint32_t)(igr*fp.nch);
            for (int ch = 0; ch < fp.nch; ch++)
            {
                const GranuleSide &g = fp.side[igr*fp.nch + ch];
                fill_gc_desc(&gcs[igr*fp.nch + ch], g, fp.hdr, ch, fp.nch, igr, bit);
                bit += g.part_23_length;
            }
        }
        for (int k = n_gc; k < 4; k++) { memset(&gcs[k], 0, sizeof(GcDesc)); gcs[k].flags1 = GCF1_MPEG1; }
        TileDesc tile;
        memset(&tile, 0, sizeof(tile));
        tile.gc_first = 0; tile.n_gr = (uint8_t)fp.n_gr; tile.nch = (uint8_t)fp.nch; tile.pcm_off = 0;
        tile.halo[0][0] = tile.halo[0][1] = tile.halo[1][0] = tile.halo[1][1] = -1;
        memcpy(h + md_bytes, gcs, sizeof(gcs));
        memcpy(h + md_bytes + sizeof(gcs), &tile, sizeof(tile));
        float *hs = (float *)ctx->h_state.p;
        memcpy(hs, dec->mdct_overlap, 576*4);
        memcpy(hs + 576, dec->qmf_state, 960*4);
        cudaStream_t s = ctx->stream;
        cudaMemcpyAsync(B.d_md.p, h, md_bytes, cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(B.d_gcs.p, h + md_bytes, sizeof(gcs), cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(B.d_tiles.p, h + md_bytes + sizeof(gcs), sizeof(tile), cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(ctx->d_state_in.p, hs, 1536*4, cudaMemcpyHostToDevice, s);
        if (n_ist) cudaMemcpyAsync(B.d_istlist.p, ist_list, (size_t)n_ist*4, cudaMemcpyHostToDevice, s);
        static const uint32_t identity_order[4] = { 0, 1, 2, 3 };
        cudaMemcpyAsync(B.d_order.p, identity_order, 16, cudaMemcpyHostToDevice, s);
        EntropyArgs ea;
        memset(&ea, 0, sizeof(ea));
        ea.md_words = (const uint32_t *)B.d_md.p; ea.md_n_words = md_bytes/4;
        ea.gcs = (const GcDesc *)B.d_gcs.p; ea.n_gc = 4;
        ea.order = (const uint32_t *)B.d_order.p; ea.n_order = (uint32_t)n_gc; ea.group_counter = (uint32_t *)B.d_counter.p;
        ea.xr = (float *)B.d_xr.p; ea.nzl = (uint16_t *)B.d_nzl.p; ea.ist_pos = (uint8_t *)B.d_istpos.p;
        cudaError_t e = launch_entropy(ctx->tables, ea, s);
        if (e == cudaSuccess && n_ist)
        {
            StereoArgs sa;
            sa.gcs = ea.gcs; sa.granules = (const uint32_t *)B.d_istlist.p; sa.n_granules = (uint32_t)n_ist;
            sa.xr = ea.xr; sa.nzl = ea.nzl; sa.ist_pos = ea.ist_pos;
            e = launch_intensity_stereo(ctx->tables, sa, s);
        }
        if (e == cudaSuccess)
        {
            TransformArgs ta;
            memset(&ta, 0, sizeof(ta));
            ta.gcs = ea.gcs; ta.tiles = (const TileDesc *)B.d_tiles.p; ta.n_tiles = 1;
            ta.xr = ea.xr; ta.nzl = ea.nzl; ta.tile_recs = (TileRec *)B.d_tile_recs.p; ta.tile_jobs = (uint16_t *)B.d_tile_jobs.p; ta.pcm = B.d_pcm.p; ta.float_output = float_output;
            ta.state_in = (const float *)ctx->d_state_in.p; ta.state_out = (float *)ctx->d_state_out.p;
            e = launch_transform(ctx->tables, ta, s);
        }
        if (e == cudaSuccess)
        {
            cudaMemcpyAsync(B.h_pcm.p, B.d_pcm.p, (size_t)n_samples*ss, cudaMemcpyDeviceToHost, s);
            cudaMemcpyAsync(hs + 1536, ctx->d_state_out.p, 1536*4, cudaMemcpyDeviceToHost, s);
            e = cudaStreamSynchronize(s);
        }
        if (e != cudaSuccess)
        {
            fprintf(stderr, "minimp3_b200: frame decode failed: %s\n", cudaGetErrorString(e));
            return 0;
        }
        memcpy(pcm, B.h_pcm.p, (size_t)n_samples*ss);
        const float *so = hs + 1536;
        for (int ch = 0; ch < fp.nch; ch++)
        {
            memcpy(dec->mdct_overlap[ch], so + ch*288, 288*4);
            memcpy(dec->qmf_state + ch*480, so + 576 + ch*480, 480*4);
        }
        result = (int)hdr_frame_samples(fp.hdr);
    }
    /* L3_save_reservoir (minimp3.h:1212-1226) */
    {
        int consumed_bits = 0;
        if (fp.success)
            for (int k = 0; k < fp.n_gr*fp.nch; k++) consumed_bits += fp.side[k].part_23_length;
        int pos = (consumed_bits + 7)/8;
        int remains = md_len - pos;
        if (remains > 511) { pos += remains - 511; remains = 511; }
        if (remains > 0) memmove(dec->reserv_buf, maindata + pos, (size_t)remains);
        dec->reserv = remains;
    }
    return result;
}

/* one frame, one GPU round trip, state in *dec (what the look-ahead layer in ex_api.cpp falls back to) */
__attribute__((visibility("hidden"))) int mp3dec_decode_frame_direct(mp3dec_t *dec, const uint8_t *mp3, int mp3_bytes, void *pcm,
                                                                     mp3dec_frame_info_t *info, int float_output)
{
    return decode_frame_impl(dec, mp3, mp3_bytes, pcm, info, float_output);
}

/* csrc/ex_api.cpp: serves the call from a look-ahead window when the caller walks a buffer frame by frame;
 * returns -1 when the call has to go through the direct path */
__attribute__((visibility("hidden"))) int mp3dec_frame_lookahead(mp3dec_t *dec, const uint8_t *mp3, int mp3_bytes, void *pcm,
                                                                 mp3dec_frame_info_t *info, int float_output);

int mp3dec_decode_frame(mp3dec_t *dec, const uint8_t *mp3, int mp3_bytes, int16_t *pcm, mp3dec_frame_info_t *info)
{
    const int n = mp3dec_frame_lookahead(dec, mp3, mp3_bytes, pcm, info, 0);
    return n >= 0 ? n : decode_frame_impl(dec, mp3, mp3_bytes, pcm, info, 0);
}

int mp3dec_decode_frame_f32(mp3dec_t *dec, const uint8_t *mp3, int mp3_bytes, float *pcm, mp3dec_frame_info_t *info)
{
    const int n = mp3dec_frame_lookahead(dec, mp3, mp3_bytes, pcm, info, 1);
    return n >= 0 ? n : decode_frame_impl(dec, mp3, mp3_bytes, pcm, info, 1);
}

}  // extern "C"

/* ---- mp3dec_f32_to_s16 (minimp3.h:1808-1864) on the device ------------------------------------------------ */

#include "l3_math.cuh"

__global__ void k_f32_to_s16(const float *__restrict__ in, int16_t *__restrict__ out, int n, int n_simd)
{
    for (int i = blockIdx.x*blockDim.x + threadIdx.x; i < n; i += gridDim.x*blockDim.x)
    {
        const float v = in[i]*32768.0f;
        out[i] = i < n_simd ? mp3b::pcm_round_simd(v) : mp3b::pcm_round_scalar(v);
    }
}

extern "C" void mp3dec_f32_to_s16(const float *in, int16_t *out, int num_samples)
{
    if (num_samples <= 0) return;
    DeviceCtx *ctx = get_ctx(0);
    if (!ctx) return;
    std::lock_guard<std::mutex> lock(ctx->mu);
    cudaSetDevice(ctx->device);
    static DevBuf d_in, d_out;
    if (!d_in.ensure((size_t)num_samples*4) || !d_out.ensure((size_t)num_samples*2)) return;
    cudaMemcpyAsync(d_in.p, in, (size_t)num_samples*4, cudaMemcpyHostToDevice, ctx->stream);
    const int blocks = std::min((num_samples + 255)/256, 148*8);
    k_f32_to_s16<<<blocks, 256, 0, ctx->stream>>>((const float *)d_in.p, (int16_t *)d_out.p, num_samples, num_samples & ~7);
    cudaMemcpyAsync(out, d_out.p, (size_t)num_samples*2, cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
}
