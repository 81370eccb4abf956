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
 * Phases (256 threads, all data in shared memory):
 *   1 load spectra (coalesced; short-block reorder folded into the store index)   [XS]
 *   2 antialias butterflies
 *   3 IMDCT: one thread per (granule-channel, subband): a) tail + own half, b) overlap with tail(g-1)  [TS, YS]
 *   4 32-point DCT-II: one thread per (channel, time slot), in place                [YS]
 *   5 window FIR: one thread per (granule, output phase j), 16 taps from registers, packed PCM store
 * Bounded by the FP32 pipe (about 60 flop per output sample) long before HBM (4 B in + 2 B out per sample).
 */
#include <cuda_runtime.h>
#include <math.h>

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
constexpr int kXPitch = 19;                 /* 18 lines + 1 pad: conflict-free per-subband access */
constexpr int kXSlot = 32*kXPitch;          /* 608 floats */
constexpr int kRows = 15 + 18*GT;           /* time slots per channel: 15 history + tile */
constexpr int kYPitch = 33;
constexpr int kXFloats = kSlots*kXSlot;     /* 12160 */
constexpr int kYFloats = 2*kRows*kYPitch;   /* 10494, aliases XS */
constexpr int kTFloats = kSlots*288;
static_assert(kYFloats <= kXFloats, "Y must fit in the X region");
static_assert(GT*32 <= kThreads, "one FIR job per thread");

__constant__ ImdctConsts c_k;

struct SlotMeta
{
    int gc;          /* granule-channel index or -1 (zero state) */
    int nzl;
    int block_type;  /* 0..3 */
    int n_long;      /* leading subbands that always use the normal long window when mixed_block_flag is
                        set -- for ANY block type, as in minimp3.h:1260 (0, 2, 4) */
    int map;         /* reorder map row or -1 */
};

struct Smem
{
    float x[kXFloats];      /* spectra, later time/subband matrix Y */
    float t[kTFloats];      /* tails [slot][i][sb] */
    SlotMeta meta[kSlots];
};

template <bool FLOAT_OUT>
__device__ __forceinline__ void store_pcm(void *pcm, uint64_t idx, float l, float r, int nch, bool scalar_round)
{
    if (FLOAT_OUT)
    {
        float *p = reinterpret_cast<float *>(pcm);
        const float k = 1.0f/32768.0f;
        if (nch == 2) *reinterpret_cast<float2 *>(p + idx) = make_float2(l*k, r*k);
        else p[idx] = l*k;
    } else
    {
        int16_t *p = reinterpret_cast<int16_t *>(pcm);
        const int16_t a = scalar_round ? pcm_round_scalar(l) : pcm_round_simd(l);
        if (nch == 2)
        {
            const int16_t b = scalar_round ? pcm_round_scalar(r) : pcm_round_simd(r);
            *reinterpret_cast<uint32_t *>(p + idx) = (uint32_t)(uint16_t)a | ((uint32_t)(uint16_t)b << 16);
        } else
            p[idx] = a;
    }
}

template <bool FLOAT_OUT>
__global__ void __launch_bounds__(kThreads, 2)
k_l3_transform(TransformArgs a, const uint16_t *__restrict__ reorder_map, const float *__restrict__ fir_coef)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TileDesc tile = a.tiles[blockIdx.x];
    const int nch = tile.nch, n_gr = tile.n_gr;
    const int n_slots = (n_gr + 2)*nch;
    const bool stateful = a.state_in != nullptr;

    /* ---- phase 0: slot table ---- */
    if (tid < kSlots)
    {
        SlotMeta m;
        m.gc = -1; m.nzl = 0; m.block_type = 0; m.n_long = 0; m.map = -1;
        if (tid < n_slots)
        {
            const int i = tid/nch - 2, ch = tid%nch;
            int gc = i >= 0 ? (int)tile.gc_first + i*nch + ch : (stateful ? -1 : tile.halo[ch][-1 - i]);
            if (gc >= 0)
            {
                const GcDesc d = a.gcs[gc];
                m.gc = gc;
                m.nzl = a.nzl[gc];
                m.block_type = d.flags0 & GCF_BLOCK_TYPE_MASK;
                const bool mixed = d.flags0 & GCF_MIXED;
                m.n_long = mixed ? (d.sr_idx == 1 ? 4 : 2) : 0;   /* 8 kHz MPEG-2.5 keeps 4 long subbands */
                if (m.block_type == 2) m.map = d.sr_idx*2 + (mixed ? 1 : 0);
            }
        }
        S.meta[tid] = m;
    }
    __syncthreads();

    /* ---- phase 1: spectra -> shared (short blocks: [window][freq] -> [freq][window] on the fly) ---- */
    for (int slot = warp; slot < n_slots; slot += kThreads/32)
    {
        const SlotMeta m = S.meta[slot];
        float *dst = S.x + slot*kXSlot;
        const float *src = m.gc >= 0 ? a.xr + (size_t)m.gc*576 : nullptr;
        const uint16_t *map = m.map >= 0 ? reorder_map + m.map*576 : nullptr;
#pragma unroll
        for (int k = 0; k < 18; k++)
        {
            const int e = lane + 32*k;
            const float v = e < m.nzl ? __ldg(src + e) : 0.f;
            const int dl = map ? (int)__ldg(map + e) : e;
            const int sb = dl/18;
            dst[sb*kXPitch + (dl - sb*18)] = v;
        }
    }
    __syncthreads();

    /* ---- phase 2: antialias (minimp3.h:1012-1045): 8 butterflies across each long/long subband boundary ---- */
    for (int slot = 0; slot < n_slots; slot++)
    {
        const SlotMeta m = S.meta[slot];
        if (m.gc < 0) continue;
        const int aa_bands = m.block_type == 2 ? m.n_long - 1 : 31;
        const int b = tid >> 3, i = tid & 7;
        if (b < aa_bands)
        {
            float *x = S.x + slot*kXSlot;
            const float u = x[(b + 1)*kXPitch + i];
            const float dn = x[b*kXPitch + 17 - i];
            x[(b + 1)*kXPitch + i] = u*c_k.aa_cs[i] - dn*c_k.aa_ca[i];
            x[b*kXPitch + 17 - i] = u*c_k.aa_ca[i] + dn*c_k.aa_cs[i];
        }
    }
    __syncthreads();

    /* ---- phase 3a: per (slot, subband): tail and own half ---- */
    constexpr int kJobs3 = (kSlots*32 + kThreads - 1)/kThreads;   /* 3 */
    float own[kJobs3][9];
#pragma unroll
    for (int r = 0; r < kJobs3; r++)
    {
        const int job = tid + r*kThreads;
        const int slot = job >> 5, sb = job & 31;
        if (slot < n_slots)
        {
            const SlotMeta m = S.meta[slot];
            float x[18], tail[9];
            const float *src = S.x + slot*kXSlot + sb*kXPitch;
#pragma unroll
            for (int k = 0; k < 18; k++) x[k] = src[k];
            if (m.block_type == 2 && sb >= m.n_long) imdct_short_core(x, own[r], tail, c_k.twid3);
            else imdct36_core(x, own[r], tail, c_k.twid9);
            float *t = S.t + slot*288 + sb;
            if (stateful && slot < 2*nch)
            {
                /* carried state replaces the halo: overlap of the previous call */
                const int ch = slot%nch;
                const bool newest = slot >= nch;
#pragma unroll
                for (int i = 0; i < 9; i++) t[i*32] = newest ? a.state_in[ch*288 + sb*9 + i] : 0.f;
            } else
            {
#pragma unroll
                for (int i = 0; i < 9; i++) t[i*32] = tail[i];
            }
        }
    }
    __syncthreads();

    /* ---- phase 3b: overlap with the previous granule's tail, frequency inversion, write Y[time][subband] ---- */
#pragma unroll
    for (int r = 0; r < kJobs3; r++)
    {
        const int job = tid + r*kThreads;
        const int slot = job >> 5, sb = job & 31;
        if (slot < n_slots && slot >= nch)
        {
            const SlotMeta m = S.meta[slot];
            const int i = slot/nch - 2, ch = slot%nch;
            float prev[9], out[18];
            const float *t = S.t + (slot - nch)*288 + sb;
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

    /* ---- phase 4: 32-point DCT-II over subbands for every time slot (minimp3.h:1274-1427) ---- */
    {
        const int rows_ch = 15 + 18*n_gr;
        const int n_jobs = rows_ch*nch;
        for (int job = tid; job < n_jobs; job += kThreads)
        {
            const int ch = job/rows_ch, row = job - ch*rows_ch;
            float *y = S.x + (ch*kRows + row)*kYPitch;
            if (stateful && row < 15)
            {
                /* carried synthesis history is stored already transformed */
                for (int k = 0; k < 32; k++) y[k] = a.state_in[576 + (ch*15 + row)*32 + k];
                continue;
            }
            float v[32];
#pragma unroll
            for (int k = 0; k < 32; k++) v[k] = y[k];
            dct2_32(v, c_k.dct_sec);
#pragma unroll
            for (int k = 0; k < 32; k++) y[k] = v[k];
        }
    }
    __syncthreads();

    /* ---- phase 5: 512-tap window as a 16-tap FIR per output phase (ISO 11172-3 fig. A.2 restated) ---- */
    {
        const int i = tid >> 5, j = tid & 31;
        if (i < n_gr)
        {
            float c[16];
#pragma unroll
            for (int k = 0; k < 16; k += 4)
            {
                const float4 q = __ldg(reinterpret_cast<const float4 *>(fir_coef + j*16 + k));
                c[k] = q.x; c[k + 1] = q.y; c[k + 2] = q.z; c[k + 3] = q.w;
            }
            /* V_t[j] = +-y[ia], V_t[32+j] = -y[ib]; signs are folded into c[] */
            const int ia = j < 16 ? 16 + j : (j == 16 ? 0 : 48 - j);
            const int ib = j <= 16 ? 16 - j : j - 16;
            const bool scalar_round = (j == 0) || (j == 16);
            float res0[18];
            for (int ch = 0; ch < nch; ch++)
            {
                const float *y = S.x + (ch*kRows + i*18)*kYPitch;
                float A[16], B[16];
#pragma unroll
                for (int h = 0; h < 15; h++)
                {
                    A[h] = y[h*kYPitch + ia];
                    B[h] = y[h*kYPitch + ib];
                }
#pragma unroll
                for (int tt = 0; tt < 18; tt++)
                {
                    const int cur = (15 + tt) & 15;
                    A[cur] = y[(15 + tt)*kYPitch + ia];
                    B[cur] = y[(15 + tt)*kYPitch + ib];
                    float acc = 0.f;
#pragma unroll
                    for (int m = 0; m < 8; m++)
                    {
                        acc = fmaf(c[2*m], A[(15 + tt - 2*m) & 15], acc);
                        acc = fmaf(c[2*m + 1], B[(15 + tt - 2*m - 1) & 15], acc);
                    }
                    if (nch == 2 && ch == 0) res0[tt] = acc;
                    else
                    {
                        const uint64_t idx = tile.pcm_off + (uint64_t)((i*576 + tt*32 + j)*nch);
                        if (nch == 2) store_pcm<FLOAT_OUT>(a.pcm, idx, res0[tt], acc, 2, scalar_round);
                        else store_pcm<FLOAT_OUT>(a.pcm, idx, acc, 0.f, 1, scalar_round);
                    }
                }
            }
        }
    }

    /* ---- carried state out (single-frame API): tails of the last granule + last 15 transformed slots ---- */
    if (a.state_out)
    {
        const int last_slot0 = (n_gr + 1)*nch;
        for (int k = tid; k < nch*288; k += kThreads)
        {
            const int ch = k/288, rem = k - ch*288, sb = rem/9, i = rem - sb*9;
            a.state_out[ch*288 + rem] = S.t[(last_slot0 + ch)*288 + i*32 + sb];
        }
        for (int k = tid; k < nch*15*32; k += kThreads)
        {
            const int ch = k/480, rem = k - ch*480, row = rem >> 5, col = rem & 31;
            a.state_out[576 + ch*480 + rem] = S.x[(ch*kRows + 18*n_gr + row)*kYPitch + col];
        }
    }
}

}  // namespace

int tables_upload(DeviceTables *t)
{
    ImdctConsts k;
    build_consts(&k);
    if (cudaMemcpyToSymbol(c_k, &k, sizeof(k)) != cudaSuccess) return -1;
    static uint16_t map[16*576];
    static float fir[32*16];
    build_reorder_map(map);
    build_fir(fir);
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
    t->reorder_map = nullptr;
    t->fir_coef = nullptr;
}

cudaError_t launch_transform(const DeviceTables &t, const TransformArgs &a, cudaStream_t s)
{
    if (!a.n_tiles) return cudaSuccess;
    const size_t smem = sizeof(Smem);
    cudaError_t e;
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

}  // namespace mp3b
