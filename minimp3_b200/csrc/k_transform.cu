/*
 * k_transform.cu -- reorder + antialias + IMDCT/overlap + frequency inversion + 32-band polyphase
 * synthesis (DCT-II matrixing and 512-tap window) + PCM conversion, one CTA per tile.
 *
 * Restates L3_reorder, L3_antialias, L3_imdct_gr, L3_change_sign (minimp3.h:995-1210) and
 * mp3d_synth_granule / mp3d_DCT_II / mp3d_synth / mp3d_synth_pair / mp3d_scale_pcm
 * (minimp3.h:1274-1655) for sm_100a.
 *
 * A tile is up to MP3B_TILE_GRANULES consecutive granules (both channels) of one stream.  The state the
 * reference carries from granule to granule (mdct_overlap, qmf_state) is re-derived from the two
 * granules in front of the tile ("halo"), so every tile is independent:
 *     tail(g)       depends on spectrum(g)                                 (9 values per subband)
 *     subband(g)    depends on spectrum(g) and tail(g-1)
 *     pcm(g)        depends on subband(g) and the last 15 time slots of subband(g-1)
 * k_l3_tile_meta (one warp per tile) first turns descriptors + decoded extents into a 256-byte tile record and the
 * list of IMDCT jobs that are not identically zero, so the transform CTA starts from one round trip.
 * Phases of k_l3_transform (256 threads, all data in shared memory):
 *   0+1 per warp, no block barrier: slot records -> one TMA bulk copy per row (decoded extent only) on the
 *       warp's mbarrier; rows cleared beyond the extent while the copies fly; plain MS stereo in place   [XS]
 *   2+3a antialias + IMDCT per listed (slot, subband) job: tail to shared memory, own half in registers  [TS]
 *   3b  overlap with tail(g-1), frequency inversion, write Y[time][subband]                              [YS]
 *   4   32-point DCT-II: one thread per (channel, time slot), in place; half-size transforms of the Lee
 *       recursion paired in packed FP32 (FFMA2)                                                          [YS]
 *   5   window FIR: one thread per (granule, output phase j), taps and history as FP32 pairs (FFMA2),
 *       packed PCM store
 * Bounded by the issue / FP32 pipes long before HBM (4 B in + 2 B out per sample).
 */
#include <cuda_runtime.h>
#include <math.h>

#include <cstdlib>
#include "kernels.h"
#include "l3_math.cuh"
#define MP3T_QUAL static const
#include "mp3_tables.h"
#include "l3_tables_build.h"

namespace mp3b {

namespace {

constexpr int GT = MP3B_TILE_GRANULES;
constexpr int kThreads = 256;
constexpr int kSlots = (GT + 2)*2;          /* granule-channel slots incl. two halo granules per channel */
constexpr int kXPitch = 18;                 /* rows sit in shared memory exactly as in HBM (bulk copies) */
constexpr int kXSlot = 588;                 /* row slot pitch: 576 lines + 12, so that the same subband of neighbouring slots falls into
                                               different banks when the IMDCT jobs of a warp span several granule-channels */
constexpr int kTSlot = 296;                 /* likewise for the tails: 288 + 8 */
constexpr int kRows = 15 + 18*GT;           /* time slots per channel: 15 history + tile */
constexpr int kYPitch = 34;                 /* even: a DCT thread moves its row as 8-byte words, 16 threads cover all banks */
constexpr int kXFloats = kSlots*kXSlot;     /* 11760 */
constexpr int kYFloats = 2*kRows*kYPitch;   /* 10812, aliases XS */
constexpr int kTFloats = kSlots*kTSlot;
static_assert(sizeof(TileRec) == 256 && MP3B_TILE_SLOTS == kSlots, "tile record layout");
static_assert(kYFloats <= kXFloats, "Y must fit in the X region");
static_assert(kXSlot % 4 == 0 && kTSlot % 4 == 0, "16-byte aligned rows (bulk-copy destinations, float4 clears)");
static_assert(GT*32 <= kThreads, "one FIR job per thread (stereo; a mono tile takes two rounds)");
static_assert(15 + 18*MP3B_TILE_GRANULES_MONO <= 2*kRows, "a mono tile's rows fit the two channels' rows");
static_assert(((GT + 2 + kThreads/32 - 1)/(kThreads/32))*2 <= 4, "a warp clears at most four rows, a quarter-warp each");
static_assert(kXFloats - kYFloats >= GT*2*18*2, "slack behind Y holds the specially rounded outputs");

__constant__ ImdctConsts c_k;
__constant__ unsigned long long c_sec2[16];   /* (s, s) pairs of dct_sec[16..30]: levels 16, 8, 4, 2 of the Lee recursion */

struct SlotMeta
{
    int gc;          /* granule-channel index or -1 (zero state) */
    int nzl;
    int block_type;  /* 0..3 */
    int n_long;      /* leading subbands that always use the normal long window when mixed_block_flag is
                        set -- for ANY block type, as in minimp3.h:1260 (0, 2, 4) */
    int map;         /* reorder map row or -1 */
    int ms;          /* plain MS stereo applied on load: 0 none, 1 this row = L + R, 2 this row = L - R */
    int act;         /* subbands that can hold non-zero lines after antialias (0..32) */
};

struct Smem
{
    alignas(128) float x[kXFloats];      /* spectra, later time/subband matrix Y */
    float t[kTFloats];      /* tails [slot][i][sb] */
    SlotMeta meta[kSlots];
    unsigned long long mbar[kThreads/32];   /* one transaction barrier per warp: its rows' bulk copies */
    alignas(16) float fir_taps[512];
};


/* ---- packed FP32 pairs (PTX f32x2, sm_100: FFMA2 / FADD2 / FMUL2) ---- */
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f32x2(unsigned long long v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma_f32x2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

__device__ __forceinline__ unsigned long long add_f32x2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long sub_f32x2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long mul_f32x2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

/* Lee recursion on pairs: the two half-size transforms every level of dct2_32 (l3_math.cuh) spawns do the same
 * operations with the same constants, so they ride in the two lanes of packed FP32 instructions.  Lane-wise
 * identical to the scalar recursion. */
template <int N>
struct Dct2Pair
{
    static __device__ __forceinline__ void run(unsigned long long *x, const unsigned long long *sec)
    {
        unsigned long long a[N/2], b[N/2];
#pragma unroll
        for (int i = 0; i < N/2; i++)
        {
            a[i] = add_f32x2(x[i], x[N - 1 - i]);
            b[i] = mul_f32x2(sub_f32x2(x[i], x[N - 1 - i]), sec[i]);
        }
        Dct2Pair<N/2>::run(a, sec + N/2);
        Dct2Pair<N/2>::run(b, sec + N/2);
#pragma unroll
        for (int k = 0; k < N/2; k++)
        {
            x[2*k] = a[k];
            x[2*k + 1] = (k + 1 < N/2) ? add_f32x2(b[k], b[k + 1]) : b[k];
        }
    }
};
template <>
struct Dct2Pair<1>
{
    static __device__ __forceinline__ void run(unsigned long long *, const unsigned long long *) {}
};

/* 32-point DCT-II of dct2_32(): first butterfly level in scalars, the two 16-point halves as one packed
 * transform, scalar recombination */
__device__ __forceinline__ void dct2_32_paired(float *x)
{
    unsigned long long p[16];
#pragma unroll
    for (int i = 0; i < 16; i++)
        p[i] = pack_f32x2(x[i] + x[31 - i], (x[i] - x[31 - i])*c_k.dct_sec[i]);
    Dct2Pair<16>::run(p, c_sec2);
    float a[16], b[16];
#pragma unroll
    for (int k = 0; k < 16; k++) unpack_f32x2(p[k], a[k], b[k]);
#pragma unroll
    for (int k = 0; k < 16; k++)
    {
        x[2*k] = a[k];
        x[2*k + 1] = k + 1 < 16 ? b[k] + b[k + 1] : b[k];
    }
}

/* ---- TMA bulk copy + transaction barrier (PTX ISA: cp.async.bulk, mbarrier) ---- */
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do
    {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

/* saturating round-to-nearest-even float -> int16 in one instruction: what the reference's
 * min/max + cvtps2dq + packssdw sequence computes (minimp3.h:1541-1542) */
__device__ __forceinline__ uint32_t f2s16_rn_sat(float x)
{
    short r;
    asm("cvt.rni.sat.s16.f32 %0, %1;" : "=h"(r) : "f"(x));
    return (uint32_t)(uint16_t)r;
}

/* the 32 x 16 window taps, staged once per CTA in shared memory as [4 chunks][32 phases][4 taps] so that the
 * FIR threads of a warp (one per phase) fetch them with conflict-free 16-byte loads */
__device__ __forceinline__ void stage_fir_taps(float *fir_taps, const float *__restrict__ fir_coef, int tid)
{
    if (tid < 128)
        reinterpret_cast<float4 *>(fir_taps)[tid] = __ldg(reinterpret_cast<const float4 *>(fir_coef) + (tid & 31)*4 + (tid >> 5));
}

/* ---- phase 5: 512-tap window as a 16-tap FIR per output phase (ISO 11172-3 fig. A.2 restated) ----
 * pcm[32 t + j] = sum_m  D[64m + j] V_{t-2m}[j] + D[64m + 32 + j] V_{t-2m-1}[32 + j];  V_t[j] = +-y_t[ia],
 * V_t[32 + j] = -y_t[ib] (signs folded into the taps).  One thread = one output phase j of one granule, both
 * channels; the 16-slot history lives in registers.  The two phases the reference rounds with its scalar
 * helper (j = 0, 16) are parked in shared memory and finished by a separate pass. */
template <bool FLOAT_OUT, int NCH, int SLOTS, int ROWS>
__device__ __forceinline__ void fir_phase(const float *ybase, const float *fir_taps, float *special,
                                          void *pcm, uint64_t pcm_off, int i, int j)
{
    /* taps as (c[2m], c[2m+1]) pairs, history as (A[h], B[h-1]) pairs: one packed FMA (Blackwell FFMA2) advances
     * the A-sum and the B-sum of the reference formula together; lanes of the pair never mix, so every partial
     * sum is the one two scalar FMA chains would produce */
    unsigned long long c2[8];
#pragma unroll
    for (int k = 0; k < 8; k += 2)
    {
        const ulonglong2 q = reinterpret_cast<const ulonglong2 *>(fir_taps)[(k >> 1)*32 + j];   /* [chunk][phase]: no bank conflicts */
        c2[k] = q.x; c2[k + 1] = q.y;
    }
    const int ia = j < 16 ? 16 + j : (j == 16 ? 0 : 48 - j);
    const int ib = j <= 16 ? 16 - j : j - 16;
    const bool is_special = !FLOAT_OUT && ((j & 15) == 0);
    float res[NCH][SLOTS];
#pragma unroll
    for (int ch = 0; ch < NCH; ch++)
    {
        const float *ya = ybase + (ch*ROWS + i*SLOTS)*kYPitch + ia;
        const float *yb = ybase + (ch*ROWS + i*SLOTS)*kYPitch + ib;
        unsigned long long P[16];
#pragma unroll
        for (int h = 1; h < 15; h++) P[h] = pack_f32x2(ya[h*kYPitch], yb[(h - 1)*kYPitch]);
        static_assert(SLOTS % 2 == 0, "two time slots per round");
#pragma unroll
        for (int tt = 0; tt < SLOTS; tt += 2)
        {
            /* two time slots per round: their FMA chains touch disjoint (even / odd) history slots and interleave */
            const int r0 = 15 + tt, r1 = 16 + tt;
            P[r0 & 15] = pack_f32x2(ya[r0*kYPitch], yb[(r0 - 1)*kYPitch]);
            P[r1 & 15] = pack_f32x2(ya[r1*kYPitch], yb[(r1 - 1)*kYPitch]);
            unsigned long long acc0 = 0ull, acc1 = 0ull;
#pragma unroll
            for (int m = 0; m < 8; m++)
            {
                acc0 = fma_f32x2(c2[m], P[(r0 - 2*m) & 15], acc0);
                acc1 = fma_f32x2(c2[m], P[(r1 - 2*m) & 15], acc1);
            }
            float lo, hi;
            unpack_f32x2(acc0, lo, hi);
            res[ch][tt] = lo + hi;
            unpack_f32x2(acc1, lo, hi);
            res[ch][tt + 1] = lo + hi;
        }
    }
    if (FLOAT_OUT)
    {
        float *p = reinterpret_cast<float *>(pcm) + pcm_off + (size_t)(i*SLOTS*32 + j)*NCH;
        const float k = 1.0f/32768.0f;
#pragma unroll
        for (int tt = 0; tt < SLOTS; tt++)
        {
            if (NCH == 2) *reinterpret_cast<float2 *>(p + tt*64) = make_float2(res[0][tt]*k, res[NCH - 1][tt]*k);
            else p[tt*32] = res[0][tt]*k;
        }
    } else if (is_special)
    {
        float *sp = special + ((i*2 + (j >> 4))*SLOTS)*NCH;
#pragma unroll
        for (int tt = 0; tt < SLOTS; tt++)
        {
            sp[tt*NCH] = res[0][tt];
            if (NCH == 2) sp[tt*NCH + 1] = res[NCH - 1][tt];
        }
    } else
    {
        int16_t *p = reinterpret_cast<int16_t *>(pcm) + pcm_off + (size_t)(i*SLOTS*32 + j)*NCH;
#pragma unroll
        for (int tt = 0; tt < SLOTS; tt++)
        {
            if (NCH == 2)
                *reinterpret_cast<uint32_t *>(p + tt*64) = f2s16_rn_sat(res[0][tt]) | (f2s16_rn_sat(res[NCH - 1][tt]) << 16);
            else
                p[tt*32] = (int16_t)f2s16_rn_sat(res[0][tt]);
        }
    }
}

/* the 2-of-32 outputs the reference produces through mp3d_scale_pcm (truncate(x + .5) - (x < 0)).  Warp i parked
 * the values of granule i itself (its lanes 0 and 16), so the pass is warp-local: no block barrier. */
template <int NCH, int SLOTS>
__device__ __forceinline__ void special_round_pass(const float *special, void *pcm, uint64_t pcm_off, int i, int lane)
{
    __syncwarp();
    for (int k = lane; k < 2*SLOTS; k += 32)
    {
        const int h = k >= SLOTS ? 1 : 0, tt = k - h*SLOTS;
        const float *sp = special + ((i*2 + h)*SLOTS + tt)*NCH;
        int16_t *p = reinterpret_cast<int16_t *>(pcm) + pcm_off + (size_t)(i*SLOTS*32 + tt*32 + h*16)*NCH;
        const int16_t l = pcm_round_scalar(sp[0]);
        if (NCH == 2)
        {
            const int16_t r = pcm_round_scalar(sp[1]);
            *reinterpret_cast<uint32_t *>(p) = (uint32_t)(uint16_t)l | ((uint32_t)(uint16_t)r << 16);
        } else
            *p = l;
    }
}

/* ---- tile records: one warp per tile (lane = slot) turns the tile descriptor, the granule descriptors and the
 * extents left by the entropy / intensity-stereo kernels into the TileRec the transform CTA starts from ---- */
__global__ void __launch_bounds__(256)
k_l3_tile_meta(TransformArgs a)
{
    const int lane = threadIdx.x & 31;
    const uint32_t t = blockIdx.x*(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= a.n_tiles) return;
    const TileDesc tile = a.tiles[t];
    const bool stateful = a.state_in != nullptr;
    const int nch = tile.nch, n_slots = (tile.n_gr + 2)*nch;
    int gc = -1, nzl = 0, act = stateful ? 32 : 0, bt = 0, ms = 0, n_long = 0, map = -1;
    if (lane < n_slots)
    {
        const int i = lane/nch - 2, ch = lane%nch;
        gc = i >= 0 ? (int)tile.gc_first + i*nch + ch : (stateful ? -1 : tile.halo[ch][-1 - i]);
        if (gc >= 0)
        {
            const GcDesc d = a.gcs[gc];
            nzl = a.nzl[gc];
            bt = d.flags0 & GCF_BLOCK_TYPE_MASK;
            const bool mixed = d.flags0 & GCF_MIXED;
            n_long = mixed ? (d.sr_idx == 1 ? 4 : 2) : 0;   /* 8 kHz MPEG-2.5 keeps 4 long subbands (minimp3.h:1260) */
            if (bt == 2) map = d.sr_idx*2 + (mixed ? 1 : 0);
            int ext = nzl;
            if (d.flags1 & GCF1_MS_PLAIN)
            {
                /* L3_midside_stereo (minimp3.h:879-909): stereo pairs sit on descriptor indices (2k, 2k+1) */
                ms = (d.flags1 & GCF1_CH) ? 2 : 1;
                ext = max(ext, (int)a.nzl[gc ^ 1]);
            }
            /* lines at or above the decoded extent are zero; antialias leaks 8 lines into the next subband */
            act = (stateful || bt == 2) ? 32 : (ext ? min(32, (ext + 17)/18 + 1) : 0);
        }
    }
    /* IMDCT job list: only subbands that are active in the granule itself or in the one before it (whose tail
     * overlaps into this one) need work, everything else is exactly zero */
    const int act_l = lane < n_slots ? act : 0;
    int prev_act = __shfl_up_sync(0xffffffffu, act_l, nch);
    if (lane < nch) prev_act = 0;
    int incl = lane < n_slots ? max(act_l, prev_act) : 0;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1)
    {
        const int v = __shfl_up_sync(0xffffffffu, incl, dd);
        if (lane >= dd) incl += v;
    }
    const int nj = lane < n_slots ? max(act_l, prev_act) : 0;
    /* plain MS stereo: the two rows of a granule sit in adjacent slots when both are part of the tile */
    const int p_gc = __shfl_xor_sync(0xffffffffu, gc, 1), p_nzl = __shfl_xor_sync(0xffffffffu, nzl, 1);
    const bool pair_ok = nch == 2 && ms && gc >= 0 && p_gc == (gc ^ 1);
    const int pair32 = pair_ok ? max(nzl, p_nzl) >> 5 : 0;
    const int far = (ms && !pair_ok) ? 1 : 0;
    TileRec *rec = a.tile_recs + t;
    if (lane < MP3B_TILE_SLOTS)
    {
        TileSlot ts;
        ts.gc = gc;
        ts.info = (uint32_t)(nzl >> 5) | (uint32_t)act << TSI_ACT_SHIFT | (uint32_t)bt << TSI_BT_SHIFT | (uint32_t)ms << TSI_MS_SHIFT |
                  (uint32_t)n_long << TSI_NLONG_SHIFT | (uint32_t)(map + 1) << TSI_MAP_SHIFT |
                  (uint32_t)pair32 << TSI_PAIR_SHIFT | (uint32_t)far << TSI_FAR_SHIFT;
        rec->slot[lane] = ts;
    }
    if (lane == 0)
    {
        rec->n_gr = tile.n_gr;
        rec->nch = tile.nch;
        rec->pcm_off = tile.pcm_off;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (lane == 0) rec->n_jobs = (uint16_t)total;
    uint16_t *jobs = a.tile_jobs + (size_t)t*MP3B_TILE_JOBS;
    for (int sl = 0; sl < n_slots; sl++)
    {
        const int start = __shfl_sync(0xffffffffu, incl - nj, sl), n = __shfl_sync(0xffffffffu, nj, sl);
        if (lane < n) jobs[start + lane] = (uint16_t)(sl << 8 | lane);
    }
}

template <bool FLOAT_OUT>
__global__ void __launch_bounds__(kThreads, 3)
k_l3_transform(TransformArgs a, const uint16_t *__restrict__ reorder_inv, const float *__restrict__ fir_coef)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TileRec *rec = a.tile_recs + blockIdx.x;
    stage_fir_taps(S.fir_taps, fir_coef, tid);     /* visible to everybody long before phase 5 (several barriers) */
    struct { uint8_t n_gr, nch; uint16_t n_jobs; uint64_t pcm_off; } tile;
    {
        const uint4 h = __ldg(reinterpret_cast<const uint4 *>(&rec->n_gr));
        tile.n_gr = (uint8_t)(h.x & 0xff);
        tile.nch = (uint8_t)((h.x >> 8) & 0xff);
        tile.n_jobs = (uint16_t)(h.x >> 16);
        tile.pcm_off = (uint64_t)h.z | (uint64_t)h.w << 32;
    }
    /* this thread's IMDCT jobs (slot << 8 | subband): fetched now, needed after the rows are in */
    constexpr int kJobs3 = (kSlots*32 + kThreads - 1)/kThreads;   /* 3 */
    uint16_t job_code[kJobs3];
#pragma unroll
    for (int r = 0; r < kJobs3; r++)
        job_code[r] = tid + r*kThreads < MP3B_TILE_JOBS ? __ldg(a.tile_jobs + (size_t)blockIdx.x*MP3B_TILE_JOBS + tid + r*kThreads) : (uint16_t)0;
    const int nch = tile.nch, n_gr = tile.n_gr;
    const int n_slots = (n_gr + 2)*nch;
    const bool stateful = a.state_in != nullptr;
    int clk_i = 0;
#define MP3B_STAMP() do { if (a.phase_clk && tid == 0 && clk_i < 16) a.phase_clk[(size_t)blockIdx.x*16 + clk_i++] = clock64(); } while (0)
    MP3B_STAMP();

    /* ---- phases 0+1, every warp on its own for the granules it owns (unit u = granule incl. halo belongs to
     * warp u % 8; lane = (unit, channel)): fetch the slot records (one round trip: k_l3_tile_meta did the
     * dependent work), one TMA bulk copy per row for its decoded extent on the warp's transaction barrier (lines
     * at or beyond nzl were never written by the entropy kernel and are zero by definition), clear the rest of
     * the rows and the tails meanwhile, wait, plain MS stereo in place.  Rows keep
     * the HBM line order; short blocks are de-interleaved when the IMDCT jobs read them.  No block-wide barrier
     * until the rows are complete. ---- */
    constexpr int kW = kThreads/32;
    constexpr int kU = (GT + 2 + kW - 1)/kW;   /* granules per warp, stereo tile (GT + 2 halo granules) */
    constexpr int kUm = (MP3B_TILE_GRANULES_MONO + 2 + kW - 1)/kW;   /* mono tile: twice the granules in the same rows */
    static_assert(kUm <= 4 && MP3B_TILE_GRANULES_MONO + 2 <= MP3B_TILE_SLOTS, "mono tile fits the slots, a quarter-warp per row");
    if (tid == 0 && blockIdx.x + 1024 < a.n_tiles)
        asm volatile("prefetch.global.L2 [%0];" :: "l"(a.tile_recs + blockIdx.x + 1024));   /* a later wave's record */
    if (tid == 32 && blockIdx.x + 1024 < a.n_tiles)
        asm volatile("prefetch.global.L2 [%0];" :: "l"(reinterpret_cast<const char *>(a.tile_recs + blockIdx.x + 1024) + 128));
    /* the slot records do not depend on the header: ask for all kU*2 a warp can own, keep those that exist */
    const int uj = lane >> 1, uch = lane & 1;          /* as if stereo; remapped below for mono */
    TileSlot ts_st, ts_mono;
    ts_st.gc = -1; ts_st.info = 0; ts_mono = ts_st;
    if (lane < kU*2 && (warp + kW*uj)*2 + uch < MP3B_TILE_SLOTS)
    {
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(&rec->slot[(warp + kW*uj)*2 + uch]));
        ts_st.gc = (int)v.x; ts_st.info = v.y;
    }
    if (lane < kUm && warp + kW*lane < MP3B_TILE_SLOTS)
    {
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(&rec->slot[warp + kW*lane]));
        ts_mono.gc = (int)v.x; ts_mono.info = v.y;
    }
    const int n_units = n_gr + 2;
    const int nsh = nch - 1;           /* nch is 1 or 2: slot / nch == slot >> nsh, slot % nch == slot & nsh (no integer division) */
    const int my_unit = nch == 2 ? warp + kW*uj : warp + kW*lane;
    const int my_ch = nch == 2 ? uch : 0;
    const int my_slot = my_unit*nch + my_ch;
    const bool has = (nch == 2 ? lane < kU*2 : lane < kUm) && my_unit < n_units;
    const TileSlot ts = nch == 2 ? ts_st : ts_mono;
    const uint32_t bar = smem_addr(&S.mbar[warp]);
    const unsigned has_mask = __ballot_sync(0xffffffffu, has);
    if (lane == 0 && has_mask)
    {
        mbar_init(bar, (uint32_t)__popc(has_mask));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    SlotMeta m;
    m.gc = -1; m.nzl = 0; m.block_type = 0; m.n_long = 0; m.map = -1; m.ms = 0; m.act = 0;
    if (has)
    {
        m.gc = ts.gc;
        m.nzl = (int)(ts.info & TSI_NZL32_MASK) << 5;
        m.act = (int)(ts.info >> TSI_ACT_SHIFT) & 63;
        m.block_type = (int)(ts.info >> TSI_BT_SHIFT) & 3;
        m.ms = (int)(ts.info >> TSI_MS_SHIFT) & 3;
        m.n_long = (int)(ts.info >> TSI_NLONG_SHIFT) & 7;
        m.map = ((int)(ts.info >> TSI_MAP_SHIFT) & 31) - 1;
        S.meta[my_slot] = m;
        const uint32_t bytes = (uint32_t)((m.nzl + 3) & ~3)*4u;
        if (m.gc >= 0 && bytes)
        {
            mbar_arrive_expect_tx(bar, bytes);
            bulk_g2s(smem_addr(S.x + my_slot*kXSlot), a.xr + (size_t)m.gc*576, bytes, bar);
        } else
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
    }
    /* while the copies are in flight: clear what the IMDCT jobs can read beyond the copied extent (own lines and
     * the eight the antialias butterflies take from the next subband) and the tails of the warp's rows */
    {
        const int my_from = has && m.gc >= 0 ? (m.nzl + 3) >> 2 : 0;
        const int my_to = has ? (min(576, (m.act + 1)*18) + 3) >> 2 : 0;
        /* four groups of eight lanes, group g on the row of lane g (a quarter-warp per 16-byte store phase) */
        const int g = lane >> 3, gl = lane & 7;
        const int from = __shfl_sync(0xffffffffu, my_from, g), to = __shfl_sync(0xffffffffu, my_to, g);
        const int slot = __shfl_sync(0xffffffffu, my_slot, g);
        if ((has_mask >> g) & 1)
        {
            float4 *zx = reinterpret_cast<float4 *>(S.x + slot*kXSlot);
            for (int k = from + gl; k < to; k += 8) zx[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 *zt = reinterpret_cast<float4 *>(S.t + slot*kTSlot);
#pragma unroll
            for (int k = 0; k < 9; k++) zt[gl + 8*k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncwarp();
    }
    MP3B_STAMP();
    if (has_mask) mbar_wait(bar, 0);
    MP3B_STAMP();
    if (nch == 2)
    {
        /* half a warp per granule: (L, R) <- (L + R, L - R) over the extent of the longer row, 16 bytes per lane */
        static_assert(kU == 2, "two granules per warp, sixteen lanes each");
        const int my_pair4 = has ? (int)((ts.info >> TSI_PAIR_SHIFT) & 31)*8 : 0;
        const int j = lane >> 4, l16 = lane & 15;
        const int lim4 = __shfl_sync(0xffffffffu, my_pair4, j*2);
        float4 *l = reinterpret_cast<float4 *>(S.x + (warp + j*kW)*2*kXSlot);
        for (int k = l16; k < lim4; k += 16)
        {
            const float4 vl = l[k], vr = l[k + kXSlot/4];
            l[k] = make_float4(vl.x + vr.x, vl.y + vr.y, vl.z + vr.z, vl.w + vr.w);
            l[k + kXSlot/4] = make_float4(vl.x - vr.x, vl.y - vr.y, vl.z - vr.z, vl.w - vr.w);
        }
        unsigned far_mask = __ballot_sync(0xffffffffu, has && ((ts.info >> TSI_FAR_SHIFT) & 1));
        while (far_mask)
        {
            /* the other channel of the granule is not part of the tile (halo across a channel-mode change):
             * fetch it from HBM */
            const int src_lane = __ffs(far_mask) - 1;
            far_mask &= far_mask - 1;
            const int f_gc = __shfl_sync(0xffffffffu, m.gc, src_lane), f_ms = __shfl_sync(0xffffffffu, m.ms, src_lane);
            const int f_nzl = __shfl_sync(0xffffffffu, m.nzl, src_lane), f_other = (int)__ldg(a.nzl + (f_gc ^ 1));
            float *x = S.x + __shfl_sync(0xffffffffu, my_slot, src_lane)*kXSlot;
            const float *oth = a.xr + (size_t)(f_gc ^ 1)*576;
            for (int e = lane; e < max(f_nzl, f_other); e += 32)
            {
                const float o = e < f_other ? __ldg(oth + e) : 0.f;
                x[e] = f_ms == 1 ? x[e] + o : o - x[e];
            }
        }
    }
    MP3B_STAMP();
    __syncthreads();
    MP3B_STAMP();

    /* ---- phases 2+3a: per active (slot, subband): antialias butterflies against both neighbours
     * (minimp3.h:1012-1045, recomputed on the fly instead of a separate in-place pass), then IMDCT tail and
     * own half ---- */
    float own[kJobs3][9];
    const int n_jobs3 = tile.n_jobs;
    int job_slot[kJobs3], job_sb[kJobs3];
#pragma unroll
    for (int r = 0; r < kJobs3; r++)
    {
        job_slot[r] = tid + r*kThreads < n_jobs3 ? job_code[r] >> 8 : -1;
        job_sb[r] = job_code[r] & 31;
    }
#pragma unroll
    for (int r = 0; r < kJobs3; r++)
    {
        const int job = tid + r*kThreads;
        const int slot = max(job_slot[r], 0), sb = job_sb[r];
#pragma unroll
        for (int k = 0; k < 9; k++) own[r][k] = 0.f;
        if (job < n_jobs3)
        {
            const SlotMeta m = S.meta[slot];
            float *t = S.t + slot*kTSlot + sb;
            if (stateful && slot < 2*nch)
            {
                /* carried state replaces the halo: overlap of the previous call */
                const int ch = slot & nsh;
                const bool newest = slot >= nch;
#pragma unroll
                for (int i = 0; i < 9; i++) t[i*32] = newest ? a.state_in[ch*288 + sb*9 + i] : 0.f;
            } else if (sb < m.act)
            {
                float x[18], tail[9];
                const float *src = S.x + slot*kXSlot + sb*kXPitch;
                const bool is_short = m.block_type == 2 && sb >= m.n_long;
                if (is_short)
                {
                    /* [window][frequency] -> [frequency][window] (L3_reorder, minimp3.h:1000-1010) */
                    const uint16_t *inv = reorder_inv + m.map*576 + sb*18;
                    const float *row = S.x + slot*kXSlot;
#pragma unroll
                    for (int k = 0; k < 18; k++) x[k] = row[__ldg(inv + k)];
                } else
                {
#pragma unroll
                    for (int k = 0; k < 18; k++) x[k] = src[k];
                }
                const int aa_bands = m.block_type == 2 ? m.n_long - 1 : 31;
                if (sb >= 1 && sb - 1 < aa_bands)
                {
                    /* boundary below: this subband holds the "u" side */
#pragma unroll
                    for (int i = 0; i < 8; i++) x[i] = x[i]*c_k.aa_cs[i] - src[-kXPitch + 17 - i]*c_k.aa_ca[i];
                }
                if (sb < aa_bands)
                {
                    /* boundary above: this subband holds the "d" side */
#pragma unroll
                    for (int i = 0; i < 8; i++) x[17 - i] = src[kXPitch + i]*c_k.aa_ca[i] + x[17 - i]*c_k.aa_cs[i];
                }
                if (is_short) imdct_short_core(x, own[r], tail, c_k.twid3);
                else imdct36_core(x, own[r], tail, c_k.twid9);
#pragma unroll
                for (int i = 0; i < 9; i++) t[i*32] = tail[i];
            }
        }
    }
    __syncthreads();
    MP3B_STAMP();

    /* the spectra are consumed; the time/subband matrix Y takes their place.  Phase 3b writes only the subbands
     * of its job list, phase 4 reads exactly those (everything else counts as zero without being cleared). */
    /* ---- phase 3b: overlap with the previous granule's tail, frequency inversion, write Y[time][subband] ---- */
#pragma unroll
    for (int r = 0; r < kJobs3; r++)
    {
        const int slot = job_slot[r], sb = job_sb[r];
        if (slot >= nch)
        {
            const SlotMeta m = S.meta[slot];
            const int i = (slot >> nsh) - 2, ch = slot & nsh;
            float prev[9], out[18];
            const float *t = S.t + (slot - nch)*kTSlot + sb;
#pragma unroll
            for (int k = 0; k < 9; k++) prev[k] = t[k*32];
            if (m.block_type == 2 && sb >= m.n_long) imdct_short_combine(prev, own[r], c_k.twid3, out);
            else imdct36_combine(prev, own[r], c_k.win[(m.block_type == 3 && sb >= m.n_long) ? 1 : 0], out);
            if (sb & 1)
            {
#pragma unroll
                for (int k = 1; k < 18; k += 2) out[k] = -out[k];
            }
            float *y = S.x + (ch*kRows)*kYPitch + sb;
            if (i >= 0)
            {
#pragma unroll
                for (int k = 0; k < 18; k++) y[(15 + i*18 + k)*kYPitch] = out[k];
                if (a.tap_sub)
                {
                    float *o = a.tap_sub + (size_t)m.gc*576 + sb*18;
#pragma unroll
                    for (int k = 0; k < 18; k++) o[k] = out[k];
                }
            } else if (!stateful)
            {
#pragma unroll
                for (int k = 3; k < 18; k++) y[(k - 3)*kYPitch] = out[k];
            }
        }
    }
    __syncthreads();
    MP3B_STAMP();

    /* ---- phase 4: 32-point DCT-II over subbands for every time slot (minimp3.h:1274-1427) ---- */
    {
        const int rows_ch = 15 + 18*n_gr;
        const int n_jobs = rows_ch*nch;
        for (int job = tid; job < n_jobs; job += kThreads)
        {
            const int ch = job >= rows_ch ? 1 : 0, row = job - ch*rows_ch;
            float *y = S.x + (ch*kRows + row)*kYPitch;
            if (stateful && row < 15)
            {
                /* carried synthesis history is stored already transformed */
                for (int k = 0; k < 32; k++) y[k] = a.state_in[576 + (ch*15 + row)*32 + k];
                continue;
            }
            /* subbands this row's granule-channel has jobs for: active in the granule or in the one before it */
            const int slot = (row < 15 ? 1 : 2 + (((row - 15)*3641) >> 16))*nch + ch;
            const int nj = max(S.meta[slot].act, S.meta[slot - nch].act);
            float v[32];
            float2 *y2 = reinterpret_cast<float2 *>(y);
#pragma unroll
            for (int k = 0; k < 32; k += 2)
            {
                const float2 q = y2[k >> 1];
                v[k] = k < nj ? q.x : 0.f;
                v[k + 1] = k + 1 < nj ? q.y : 0.f;
            }
            dct2_32_paired(v);
#pragma unroll
            for (int k = 0; k < 32; k += 2) y2[k >> 1] = make_float2(v[k], v[k + 1]);
        }
    }
    __syncthreads();
    MP3B_STAMP();

    /* ---- phase 5: window FIR + PCM conversion ---- */
    float *special = S.x + kYFloats;     /* slack behind Y: n_gr*2*18*nch floats */
    if (nch == 2)
    {
        if (warp < n_gr)
        {
            fir_phase<FLOAT_OUT, 2, 18, kRows>(S.x, S.fir_taps, special, a.pcm, tile.pcm_off, warp, lane);
            if (!FLOAT_OUT) special_round_pass<2, 18>(special, a.pcm, tile.pcm_off, warp, lane);
        }
    } else
    {
        /* a mono tile holds twice the granules (same rows of shared memory): a warp takes granules w, w + 8 */
        for (int i = warp; i < n_gr; i += kThreads/32)
        {
            fir_phase<FLOAT_OUT, 1, 18, kRows>(S.x, S.fir_taps, special, a.pcm, tile.pcm_off, i, lane);
            if (!FLOAT_OUT) special_round_pass<1, 18>(special, a.pcm, tile.pcm_off, i, lane);
        }
    }

    MP3B_STAMP();
    /* ---- carried state out (single-frame API): tails of the last granule + last 15 transformed slots ---- */
    if (a.state_out)
    {
        const int last_slot0 = (n_gr + 1)*nch;
        for (int k = tid; k < nch*288; k += kThreads)
        {
            const int ch = k/288, rem = k - ch*288, sb = rem/9, i = rem - sb*9;
            a.state_out[ch*288 + rem] = S.t[(last_slot0 + ch)*kTSlot + i*32 + sb];
        }
        for (int k = tid; k < nch*15*32; k += kThreads)
        {
            const int ch = k/480, rem = k - ch*480, row = rem >> 5, col = rem & 31;
            a.state_out[576 + ch*480 + rem] = S.x[(ch*kRows + 18*n_gr + row)*kYPitch + col];
        }
    }
}

/* ---- Layer I/II synthesis (mp3d_synth_granule with 12 slots, minimp3.h:1793): the blocks already hold subband
 * samples, so a tile is just: load rows (15 history slots from the two previous blocks), DCT-II, window. ---- */
constexpr int kRows12 = 15 + 12*GT;
constexpr int kY12Floats = 2*kRows12*kYPitch;

template <bool FLOAT_OUT>
__global__ void __launch_bounds__(kThreads, 3)
k_l12_synth(Synth12Args a, const float *__restrict__ fir_coef)
{
    __shared__ __align__(16) float y[kY12Floats + GT*2*12*2];
    __shared__ __align__(16) float fir_taps[512];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    stage_fir_taps(fir_taps, fir_coef, tid);
    const TileDesc tile = a.tiles[blockIdx.x];
    const int nch = tile.nch, n_gr = tile.n_gr;
    const bool stateful = a.state_in != nullptr;
    const int rows_ch = 15 + 12*n_gr;

    /* rows -> shared: row r of channel ch is slot (r - 15) of the tile; r < 15 comes from the halo blocks.  A thread
     * moves 16 bytes of a row, the rounds are independent and unrolled: four loads in flight per thread */
    const int n_quads = rows_ch*nch*8;
#pragma unroll 4
    for (int job = tid; job < n_quads; job += kThreads)
    {
        const int q = job & 7, rr = job >> 3;
        const int ch = rr >= rows_ch ? 1 : 0, r = rr - ch*rows_ch;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r >= 15)
        {
            const int b = (r - 15)/12, slot = (r - 15) - b*12;
            v = __ldg(reinterpret_cast<const float4 *>(a.sub12 + (size_t)(tile.gc_first + b*nch + ch)*384 + slot*32) + q);
        } else if (!stateful)     /* carried history is already transformed; filled after the DCT */
        {
            const int blk = r >= 3 ? tile.halo[ch][0] : tile.halo[ch][1];
            const int slot = r >= 3 ? r - 3 : 9 + r;
            if (blk >= 0) v = __ldg(reinterpret_cast<const float4 *>(a.sub12 + (size_t)blk*384 + slot*32) + q);
        }
        float2 *dst = reinterpret_cast<float2 *>(y + (ch*kRows12 + r)*kYPitch + q*4);
        dst[0] = make_float2(v.x, v.y);
        dst[1] = make_float2(v.z, v.w);
    }
    __syncthreads();
    for (int job = tid; job < rows_ch*nch; job += kThreads)
    {
        const int ch = job >= rows_ch ? 1 : 0, r = job - ch*rows_ch;
        float *row = y + (ch*kRows12 + r)*kYPitch;
        if (stateful && r < 15)
        {
            for (int k = 0; k < 32; k++) row[k] = a.state_in[(ch*15 + r)*32 + k];
            continue;
        }
        float2 *row2 = reinterpret_cast<float2 *>(row);
        float v[32];
#pragma unroll
        for (int k = 0; k < 32; k += 2) { const float2 p = row2[k >> 1]; v[k] = p.x; v[k + 1] = p.y; }
        dct2_32_paired(v);
#pragma unroll
        for (int k = 0; k < 32; k += 2) row2[k >> 1] = make_float2(v[k], v[k + 1]);
    }
    __syncthreads();
    float *special = y + kY12Floats;
    if (warp < n_gr)
    {
        if (nch == 2) fir_phase<FLOAT_OUT, 2, 12, kRows12>(y, fir_taps, special, a.pcm, tile.pcm_off, warp, lane);
        else fir_phase<FLOAT_OUT, 1, 12, kRows12>(y, fir_taps, special, a.pcm, tile.pcm_off, warp, lane);
        if (!FLOAT_OUT)
        {
            if (nch == 2) special_round_pass<2, 12>(special, a.pcm, tile.pcm_off, warp, lane);
            else special_round_pass<1, 12>(special, a.pcm, tile.pcm_off, warp, lane);
        }
    }
    if (a.state_out)
        for (int k = tid; k < nch*15*32; k += kThreads)
        {
            const int ch = k/480, rem = k - ch*480, row = rem >> 5, col = rem & 31;
            a.state_out[ch*480 + rem] = y[(ch*kRows12 + 12*n_gr + row)*kYPitch + col];
        }
}

}  // namespace

int tables_upload(DeviceTables *t)
{
    ImdctConsts k;
    build_consts(&k);
    if (cudaMemcpyToSymbol(c_k, &k, sizeof(k)) != cudaSuccess) return -1;
    {
        float sec2[32];
        for (int i = 0; i < 15; i++) sec2[2*i] = sec2[2*i + 1] = k.dct_sec[16 + i];
        sec2[30] = sec2[31] = 0.f;
        if (cudaMemcpyToSymbol(c_sec2, sec2, sizeof(sec2)) != cudaSuccess) return -1;
    }
    static uint16_t map[16*576];
    static float fir[32*16];
    {
        /* the kernel gathers: it wants the source line of every destination line */
        static uint16_t fwd[16*576];
        build_reorder_map(fwd);
        for (int r = 0; r < 16; r++)
            for (int e = 0; e < 576; e++) map[r*576 + fwd[r*576 + e]] = (uint16_t)e;
    }
    build_fir(fir);
    static uint8_t band[3*8*576];
    for (int kind = 0; kind < 3; kind++)
        for (int sr = 0; sr < 8; sr++)
        {
            const uint8_t *w = kind == 0 ? MP3T_SFB_LONG + sr*23 : kind == 1 ? MP3T_SFB_MIXED + sr*40 : MP3T_SFB_SHORT + sr*40;
            uint8_t *o = band + (kind*8 + sr)*576;
            int line = 0, b = 0;
            for (; w[b] && line < 576; b++)
                for (int k = 0; k < w[b] && line < 576; k++) o[line++] = (uint8_t)b;
            for (; line < 576; line++) o[line] = (uint8_t)(b ? b - 1 : 0);
        }
    if (cudaMalloc(&t->band_of_line, sizeof(band)) != cudaSuccess) return -1;
    if (cudaMemcpy(t->band_of_line, band, sizeof(band), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    if (cudaMalloc(&t->reorder_map, sizeof(map)) != cudaSuccess) return -1;
    if (cudaMalloc(&t->fir_coef, sizeof(fir)) != cudaSuccess) return -1;
    if (cudaMemcpy(t->reorder_map, map, sizeof(map), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    if (cudaMemcpy(t->fir_coef, fir, sizeof(fir), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    return 0;
}

void tables_free(DeviceTables *t)
{
    cudaFree(t->reorder_map);
    cudaFree(t->fir_coef);
    cudaFree(t->band_of_line);
    t->band_of_line = nullptr;
    t->reorder_map = nullptr;
    t->fir_coef = nullptr;
}

cudaError_t launch_transform(const DeviceTables &t, const TransformArgs &a, cudaStream_t s)
{
    if (!a.n_tiles) return cudaSuccess;
    /* MP3B_DEBUG_SMEM_PAD=<bytes> (read once): pads the dynamic shared memory to lower the CTAs per SM; used with
     * tools/phase_times.py to time a tile that has the SM to itself */
    static const size_t debug_pad = [] { const char *p = getenv("MP3B_DEBUG_SMEM_PAD"); return p ? (size_t)atoi(p) : (size_t)0; }();
    const size_t smem = sizeof(Smem) + debug_pad;
    cudaError_t e;
    k_l3_tile_meta<<<(a.n_tiles + 7)/8, 256, 0, s>>>(a);
    if (a.float_output)
    {
        e = cudaFuncSetAttribute(k_l3_transform<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_l3_transform<true><<<a.n_tiles, kThreads, smem, s>>>(a, t.reorder_map, t.fir_coef);
    } else
    {
        e = cudaFuncSetAttribute(k_l3_transform<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_l3_transform<false><<<a.n_tiles, kThreads, smem, s>>>(a, t.reorder_map, t.fir_coef);
    }
    return cudaGetLastError();
}

cudaError_t launch_l12_synth(const DeviceTables &t, const Synth12Args &a, cudaStream_t s)
{
    if (!a.n_tiles) return cudaSuccess;
    if (a.float_output) k_l12_synth<true><<<a.n_tiles, kThreads, 0, s>>>(a, t.fir_coef);
    else k_l12_synth<false><<<a.n_tiles, kThreads, 0, s>>>(a, t.fir_coef);
    return cudaGetLastError();
}

}  // namespace mp3b
