/*
 * k_entropy.cu -- scalefactor decode + Huffman decode + requantisation.
 *
 * Restates L3_decode_scalefactors / L3_read_scalefactors (minimp3.h:609-714) and L3_huffman
 * (minimp3.h:742-877) for sm_100a; the per-granule state machine lives in l3_entropy.cuh.  (MS stereo is
 * applied by the transform kernel while it loads the spectra, intensity stereo by k_stereo.cu.)
 *
 * Mapping: one LANE per granule-channel; a warp takes 32 granule-channels of similar main-data size (the host
 * sorts them, see build_order in host_parse.cpp) so that its lanes finish together.  Huffman decode
 * is serial inside a granule, so the parallelism is across granules; to keep the memory side coalesced
 * the warp runs a SIMT-uniform loop over line pairs (iteration k produces lines 2k, 2k+1 of all 32
 * granule-channels; a count1 quad takes two iterations) and every lane stores its pair straight into its row:
 * eight lines = one 32-byte sector are collected in registers and leave as two 16-byte stores, so
 * L2 and HBM see whole sectors and no shared-memory transpose is needed.  Code books, |q|^(4/3) and band widths sit in
 * shared memory.  Rows are only written up to the last decoded line (rounded up to 32); nzl[] tells the consumer.
 */
#include <cuda_runtime.h>
#include <stddef.h>

#include "kernels.h"
#define MP3T_QUAL static __device__ const
#include "mp3_tables.h"

namespace mp3b {
namespace {

constexpr int kWarpsPerBlock = 16;
constexpr int kBlocksPerSm = 2;      /* 12 x 3 (36 warps per SM fit at 54 registers) measured twice: within +-5 % either way depending on the stream, slower on config 2 */
constexpr int kC1Base = MP3T_HUFF_LUT_SIZE;          /* count1 table A (64 entries) then B (16) */
constexpr int kSfbShort = 8*23, kSfbMixed = 8*23 + 8*40;

struct SharedTables
{
    uint16_t lut[(MP3T_HUFF_LUT_SIZE + 80 + 7) & ~7];
    float pow43[148];
    uint32_t tabinfo[32];
    uint8_t sfb[8*23 + 8*40 + 8*40];    /* long, short, mixed band widths, one row per sample rate */
};

struct WarpScratch
{
    uint8_t prev[40*32];    /* the previous granule's scalefactors during set-up (scfsi) */
    uint8_t iscf[40*32];
};

/* the tables: static shared memory, named directly by the decoder's accessors (see l3_entropy.cuh) */
__shared__ SharedTables g_tabs;
/* the per-warp scalefactor scratch: dynamic shared memory */
extern __shared__ __align__(16) unsigned char g_scratch[];

}  // namespace
}  // namespace mp3b

/* the two tables of the decode loop are read with explicit shared-space loads from a 32-bit address: through the
 * named array the compiler rebuilt the shared window base (S2R SR_CgaCtaId, MOV, LEA) in front of every access --
 * three times per decoded pair -- rather than spend a register on it */
namespace mp3b {
namespace {
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds_u16(uint32_t a)
{
    unsigned short v;
    asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return v;
}
struct SharedF32
{
    uint32_t base;
    __device__ __forceinline__ float operator[](int i) const
    {
        float v;
        asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + 4u*(uint32_t)i));
        return v;
    }
};
}  // namespace
}  // namespace mp3b
#if defined(__CUDA_ARCH__)
#define MP3B_ENT_LUT(T, i) (mp3b::lds_u16((T).smem_base + (uint32_t)offsetof(mp3b::SharedTables, lut) + 2u*(uint32_t)(i)))
#define MP3B_ENT_POW43(T) (mp3b::SharedF32{ (T).smem_base + (uint32_t)offsetof(mp3b::SharedTables, pow43) })
#define MP3B_ENT_TABINFO(T, i) (g_tabs.tabinfo[i])
#define MP3B_ENT_SFB(T, kind, row) \
    (g_tabs.sfb + ((kind) == 0 ? (row)*23 : (kind) == 1 ? kSfbMixed + (row)*40 : kSfbShort + (row)*40))
#endif
#include "l3_entropy.cuh"

namespace mp3b {

namespace {

__device__ __forceinline__ GcDesc load_desc(const GcDesc *p)
{
    GcDesc d;
    const uint4 *src = reinterpret_cast<const uint4 *>(p);
    reinterpret_cast<uint4 *>(&d)[0] = __ldg(src);
    reinterpret_cast<uint4 *>(&d)[1] = __ldg(src + 1);
    return d;
}

template <bool TAPS>
__global__ void __launch_bounds__(kWarpsPerBlock*32, kBlocksPerSm) k_l3_entropy(EntropyArgs a)
{
    SharedTables &S = g_tabs;
    WarpScratch *scratch = reinterpret_cast<WarpScratch *>(g_scratch);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    /* ---- tables to shared memory ---- */
    for (int i = tid; i < MP3T_HUFF_LUT_SIZE; i += blockDim.x) S.lut[i] = MP3T_HUFF_LUT[i];
    for (int i = tid; i < 145; i += blockDim.x) S.pow43[i] = MP3T_POW43[i];
    if (tid < 32) S.tabinfo[tid] = (uint32_t)MP3T_HUFF_BASE[tid] | ((uint32_t)MP3T_HUFF_ROOTW[tid] << 16) | ((uint32_t)MP3T_HUFF_LINBITS[tid] << 20);
    if (tid < 64) S.lut[kC1Base + tid] = (uint16_t)(((MP3T_COUNT1A[tid] >> 4) << 8) | (MP3T_COUNT1A[tid] & 15));
    if (tid < 16) S.lut[kC1Base + 64 + tid] = (uint16_t)((4 << 8) | (~tid & 15));
    for (int i = tid; i < 8*23; i += blockDim.x) S.sfb[i] = MP3T_SFB_LONG[i];
    for (int i = tid; i < 8*40; i += blockDim.x) { S.sfb[kSfbShort + i] = MP3T_SFB_SHORT[i]; S.sfb[kSfbMixed + i] = MP3T_SFB_MIXED[i]; }
    __syncthreads();

    /* the decoder reaches the tables through the MP3B_ENT_* accessors above; the pointers only serve the host build */
    EntropyTables T;
    T.lut = S.lut; T.pow43 = S.pow43; T.tabinfo = S.tabinfo;
    T.c1info[0] = (uint32_t)kC1Base | (6u << 16);
    T.c1info[1] = (uint32_t)(kC1Base + 64) | (4u << 16);
    T.sfb_long = S.sfb; T.sfb_short = S.sfb + kSfbShort; T.sfb_mixed = S.sfb + kSfbMixed;
    /* kept in a register: coming out of an asm statement the compiler cannot rebuild it at every use */
    T.smem_base = smem_u32(&g_tabs);
    asm volatile("" : "+r"(T.smem_base));

    WarpScratch &W = scratch[warp];
    const uint32_t n_groups = (a.n_order + 31)/32;

    /* the first group of a warp is fixed, further ones come from a global counter: SMs differ in speed by ~10 %
     * and groups in length, a static split would wait for the slowest warp.  The claim is issued before the group
     * is decoded, so its round trip is free. */
    const uint32_t n_warps = gridDim.x*kWarpsPerBlock;
    for (uint32_t group = blockIdx.x*kWarpsPerBlock + warp, claimed = 0; group < n_groups; group = claimed)
    {
        if (lane == 0) claimed = atomicAdd(a.group_counter, 1u) + n_warps;
        const uint32_t item = group*32 + lane;
        const bool valid = item < a.n_order;
        const uint32_t gc = valid ? __ldg(a.order + item) : 0u;
        GcDesc d;
        if (valid) d = load_desc(a.gcs + gc);
        else
        {
            const uint4 z = make_uint4(0, 0, 0, 0);
            reinterpret_cast<uint4 *>(&d)[0] = z;
            reinterpret_cast<uint4 *>(&d)[1] = z;
            d.flags1 = GCF1_MPEG1;
        }
        const int n_sfb = d.n_long_sfb + d.n_short_sfb;

        GranuleDecoder dec;
        {
            const bool want_ist = valid && a.ist_pos && (((d.flags0 & GCF_I_STEREO) && (d.flags1 & GCF1_CH)) || a.tap_scf);
            uint8_t ist_local[40];
            int n_ist = 0;
            GcDesc d0;
            const bool use_d0 = valid && (d.flags1 & GCF1_GR1) && d.scfsi;
            if (use_d0) d0 = load_desc(a.gcs + (gc - ((d.flags1 & GCF1_STEREO) ? 2 : 1)));
            dec.init(d, use_d0 ? &d0 : nullptr, a.md_words, a.md_n_words, T, W.iscf + lane, W.prev + lane, 32,
                     want_ist ? ist_local : nullptr, &n_ist);
            if (want_ist)
            {
                uint8_t *o = a.ist_pos + (size_t)gc*40;
#pragma unroll 1
                for (int k = 0; k < 40; k++) o[k] = k < n_ist ? ist_local[k] : (uint8_t)0;
            }
        }
        if (!valid) dec.mode = EMODE_DONE;
        if (a.tap_scf && valid)
        {
            float *o = a.tap_scf + (size_t)gc*40;
            for (int i = 0; i < 40; i++) o[i] = i < n_sfb ? dec.band_scale(i) : 0.f;
        }
        int16_t *tq = (a.tap_q && valid) ? a.tap_q + (size_t)gc*576 : nullptr;

        float *row = a.xr + (size_t)gc*576;
        int iter = 0;
        float4 held = make_float4(0.f, 0.f, 0.f, 0.f);
        for (;;)
        {
            /* two steps = four lines = half a 32-byte sector */
            float4 v;
            int q0 = 0, q1 = 0, q2 = 0, q3 = 0;
            dec.template step<TAPS>(T, v.x, v.y, q0, q1);
            dec.template step<TAPS>(T, v.z, v.w, q2, q3);
            if (TAPS && tq && (q0 | q1 | q2 | q3))
            {
                tq[2*iter] = (int16_t)q0; tq[2*iter + 1] = (int16_t)q1;
                tq[2*iter + 2] = (int16_t)q2; tq[2*iter + 3] = (int16_t)q3;
            }
            /* a finished lane keeps writing zeros up to the 32-line boundary behind its last decoded line */
            const int my_end = dec.mode == EMODE_DONE ? ((dec.extent + 31) & ~31) : 576;
            if (iter & 2)
            {
                /* every other round: the 32-byte sector leaves as a whole */
                if (valid && 2*iter < my_end)
                {
                    /* streaming stores: 0.5 GB of spectra pass through, the 135 MB of bitstream should keep the L2 */
                    __stcs(reinterpret_cast<float4 *>(row + 2*iter - 4), held);
                    __stcs(reinterpret_cast<float4 *>(row + 2*iter), v);
                }
            } else
                held = v;
            iter += 2;
            bool finished = iter == 288;
            if ((iter & 15) == 0 && !finished) finished = __ballot_sync(0xffffffffu, dec.mode != EMODE_DONE) == 0;
            if (finished) break;
        }
        /* ---- per-row extent for the consumer ---- */
        if (valid) a.nzl[gc] = (uint16_t)((dec.extent + 31) & ~31);
        claimed = __shfl_sync(0xffffffffu, claimed, 0);
    }
}

}  // namespace

cudaError_t launch_entropy(const DeviceTables &, const EntropyArgs &a, cudaStream_t s)
{
    if (!a.n_order) return cudaSuccess;
    const size_t smem = sizeof(WarpScratch)*kWarpsPerBlock;     /* the tables are static shared memory */
    const bool taps = a.tap_q != nullptr;       /* the parity tests want the signed integers as well */
    cudaError_t e = taps ? cudaFuncSetAttribute(k_l3_entropy<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                         : cudaFuncSetAttribute(k_l3_entropy<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t n_groups = (a.n_order + 31)/32;
    uint32_t blocks = (n_groups + kWarpsPerBlock - 1)/kWarpsPerBlock;
    const uint32_t max_blocks = (uint32_t)sms*kBlocksPerSm;
    if (blocks > max_blocks) blocks = max_blocks;
    if (cudaMemsetAsync(a.group_counter, 0, sizeof(uint32_t), s) != cudaSuccess) return cudaGetLastError();
    if (taps) k_l3_entropy<true><<<blocks, kWarpsPerBlock*32, smem, s>>>(a);
    else k_l3_entropy<false><<<blocks, kWarpsPerBlock*32, smem, s>>>(a);
    return cudaGetLastError();
}

}  // namespace mp3b

/* ---- test tap (include/minimp3_b200.h): the requantiser's |x|^(4/3) evaluated on the device ---- */
namespace {
__global__ void k_pow43_tap(float *out, int n)
{
    const int x = blockIdx.x*blockDim.x + threadIdx.x;
    if (x < n) out[x] = mp3b::l3_pow43(x, MP3T_POW43);
}
}  // namespace

extern "C" int mp3dec_b200_pow43_device(float *dst, int n)
{
    if (!dst || n < 0 || n > 8207) return -1;     /* MP3D_E_PARAM */
    if (!n) return 0;
    float *d = nullptr;
    if (cudaMalloc(&d, (size_t)n*sizeof(float)) != cudaSuccess) return -2;    /* MP3D_E_MEMORY */
    k_pow43_tap<<<(n + 255)/256, 256>>>(d, n);
    const cudaError_t e = cudaMemcpy(dst, d, (size_t)n*sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return e == cudaSuccess ? 0 : -2;
}
