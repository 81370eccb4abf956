/*
 * k_entropy.cu -- scalefactor decode + Huffman decode + requantisation (+ plain MS stereo).
 *
 * Restates L3_decode_scalefactors / L3_read_scalefactors (minimp3.h:609-714), L3_huffman
 * (minimp3.h:742-877) and L3_midside_stereo (minimp3.h:879-909) for sm_100a.
 *
 * Mapping: one LANE per granule-channel, one warp per 32 consecutive granule-channels.  Huffman decode
 * is serial inside a granule, so the parallelism is across granules; to keep the memory side coalesced
 * the warp runs a SIMT-uniform loop over line pairs (iteration k produces lines 2k, 2k+1 of all 32
 * granule-channels; a count1 quad takes two iterations), stages 64 lines x 32 rows in shared memory and
 * flushes them as full 128-byte row segments.  Code books, |q|^(4/3) and band widths sit in shared memory.
 * Rows are only written up to the last decoded line (rounded up to 32); nzl[] tells the consumer.
 */
#include <cuda_runtime.h>

#include "kernels.h"
#include "l3_math.cuh"
#define MP3T_QUAL static __device__ const
#include "mp3_tables.h"

namespace mp3b {

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kStageLines = 64;
constexpr int kStagePitch = 33;

/* scalefactor partition sizes; kind: 0 long, 1 mixed, 2 short (ISO 11172-3 2.4.2.7, 13818-3 2.4.3.2) */
__constant__ uint8_t c_part_mpeg1[3][4] = { { 6, 5, 5, 5 }, { 8, 9, 6, 12 }, { 9, 9, 6, 12 } };
__constant__ uint8_t c_part_lsf[6][3][4] = {
    { { 6, 5, 5, 5 }, { 6, 9, 9, 9 }, { 9, 9, 9, 9 } },
    { { 6, 5, 7, 3 }, { 6, 9, 12, 6 }, { 9, 9, 12, 6 } },
    { { 11, 10, 0, 0 }, { 15, 18, 0, 0 }, { 18, 18, 0, 0 } },
    { { 7, 7, 7, 0 }, { 6, 15, 12, 0 }, { 12, 12, 12, 0 } },
    { { 6, 6, 6, 3 }, { 6, 12, 9, 6 }, { 12, 9, 9, 6 } },
    { { 8, 8, 5, 0 }, { 6, 18, 9, 0 }, { 15, 12, 9, 0 } } };
__constant__ uint8_t c_slen1[16] = { 0, 0, 0, 0, 3, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4 };
__constant__ uint8_t c_slen2[16] = { 0, 1, 2, 3, 0, 1, 2, 3, 1, 2, 3, 1, 2, 3, 2, 3 };
__constant__ uint8_t c_preamp[10] = { 1, 1, 1, 1, 2, 2, 3, 3, 3, 2 };

struct SharedTables
{
    uint16_t lut[(MP3T_HUFF_LUT_SIZE + 7) & ~7];
    float pow43[148];
    uint32_t tabinfo[32];          /* base | rootw << 16 | linbits << 20 */
    uint8_t count1a[64];
    uint8_t sfb_long[8*23];
    uint8_t sfb_short[8*40];
    uint8_t sfb_mixed[8*40];
};

struct WarpScratch
{
    float stage[kStageLines*kStagePitch];
    uint8_t iscf[40*32];
    uint8_t prev[40*32];
};

/* MSB-first bit reader over big-endian bytes, fed by aligned 32-bit loads (minimp3.h:765-768 restated) */
struct BitRd
{
    const uint32_t *w;
    uint64_t n_words, widx;
    uint64_t buf;
    int avail;
    int pos;   /* bits consumed since init */

    __device__ __forceinline__ uint32_t word(uint64_t i) const
    {
        return i < n_words ? __byte_perm(__ldg(w + i), 0, 0x0123) : 0u;
    }
    __device__ __forceinline__ void init(const uint32_t *words, uint64_t nw, uint64_t bit_off)
    {
        w = words; n_words = nw;
        widx = bit_off >> 5;
        const int sh = (int)(bit_off & 31);
        buf = ((uint64_t)word(widx) << 32) | word(widx + 1);
        buf <<= sh;
        avail = 64 - sh;
        widx += 2;
        pos = 0;
    }
    __device__ __forceinline__ void refill()
    {
        if (avail <= 32)
        {
            buf |= (uint64_t)word(widx++) << (32 - avail);
            avail += 32;
        }
    }
    __device__ __forceinline__ uint32_t peek(int n) const { return n ? (uint32_t)(buf >> (64 - n)) : 0u; }
    __device__ __forceinline__ void skip(int n) { buf <<= n; avail -= n; pos += n; }
    __device__ __forceinline__ uint32_t get(int n)
    {
        refill();
        const uint32_t v = peek(n);
        skip(n);
        return v;
    }
};

struct ScfLayout
{
    uint8_t size[4];
    uint8_t count[4];
};

/* slen / partition selection (minimp3.h:666-687) */
__device__ __forceinline__ ScfLayout scf_layout(const GcDesc &d)
{
    ScfLayout L;
    const int kind = d.n_short_sfb ? (d.n_long_sfb ? 1 : 2) : 0;
    if (d.flags1 & GCF1_MPEG1)
    {
        const int s1 = c_slen1[d.scalefac_compress & 15], s2 = c_slen2[d.scalefac_compress & 15];
        L.size[0] = L.size[1] = (uint8_t)s1;
        L.size[2] = L.size[3] = (uint8_t)s2;
#pragma unroll
        for (int i = 0; i < 4; i++) L.count[i] = c_part_mpeg1[kind][i];
    } else
    {
        const bool ist = (d.flags0 & GCF_I_STEREO) && (d.flags1 & GCF1_CH);
        int sfc = d.scalefac_compress >> (ist ? 1 : 0);
        int tn;
        if (!ist)
        {
            if (sfc < 400)      { tn = 0; L.size[0] = (sfc >> 4)/5; L.size[1] = (sfc >> 4)%5; L.size[2] = (sfc & 15) >> 2; L.size[3] = sfc & 3; }
            else if (sfc < 500) { tn = 1; sfc -= 400; L.size[0] = (sfc >> 2)/5; L.size[1] = (sfc >> 2)%5; L.size[2] = sfc & 3; L.size[3] = 0; }
            else                { tn = 2; sfc -= 500; L.size[0] = sfc/3; L.size[1] = sfc%3; L.size[2] = 0; L.size[3] = 0; }
        } else
        {
            if (sfc < 180)      { tn = 3; L.size[0] = sfc/36; L.size[1] = (sfc%36)/6; L.size[2] = sfc%6; L.size[3] = 0; }
            else if (sfc < 244) { tn = 4; sfc -= 180; L.size[0] = (sfc & 63) >> 4; L.size[1] = (sfc & 15) >> 2; L.size[2] = sfc & 3; L.size[3] = 0; }
            else                { tn = 5; sfc -= 244; L.size[0] = sfc/3; L.size[1] = sfc%3; L.size[2] = 0; L.size[3] = 0; }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) L.count[i] = c_part_lsf[tn][kind][i];
    }
    return L;
}

/* L3_read_scalefactors (minimp3.h:609-640).  dst/prev are [band][lane] byte columns in shared memory.
 * prev == nullptr means "no previous granule": copied partitions read as zero. */
__device__ __forceinline__ int read_scalefactors(BitRd &br, const ScfLayout &L, int scfsi, bool lsf, uint8_t *dst,
                                                 const uint8_t *prev, uint8_t *ist_out, int lane)
{
    int idx = 0;
    for (int p = 0; p < 4 && L.count[p]; p++, scfsi *= 2)
    {
        const int cnt = L.count[p];
        if (scfsi & 8)
        {
            for (int k = 0; k < cnt; k++)
            {
                const uint8_t v = prev ? prev[(idx + k)*32 + lane] : (uint8_t)0;
                dst[(idx + k)*32 + lane] = v;
                if (ist_out) ist_out[idx + k] = v;
            }
        } else
        {
            const int bits = L.size[p];
            if (!bits)
            {
                for (int k = 0; k < cnt; k++)
                {
                    dst[(idx + k)*32 + lane] = 0;
                    if (ist_out) ist_out[idx + k] = 0;
                }
            } else
            {
                const int max_scf = lsf ? (1 << bits) - 1 : -1;
                for (int k = 0; k < cnt; k++)
                {
                    const int s = (int)br.get(bits);
                    dst[(idx + k)*32 + lane] = (uint8_t)s;
                    if (ist_out) ist_out[idx + k] = (uint8_t)(s == max_scf ? 255 : s);
                }
            }
        }
        idx += cnt;
    }
    dst[(idx + 0)*32 + lane] = 0;
    dst[(idx + 1)*32 + lane] = 0;
    dst[(idx + 2)*32 + lane] = 0;
    return idx;
}

enum { MODE_BIG = 0, MODE_C1A = 1, MODE_C1B = 2, MODE_DONE = 3 };

__global__ void __launch_bounds__(kWarpsPerBlock*32, 2)
k_l3_entropy(EntropyArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SharedTables &T = *reinterpret_cast<SharedTables *>(smem_raw);
    WarpScratch *scratch = reinterpret_cast<WarpScratch *>(smem_raw + ((sizeof(SharedTables) + 15) & ~15));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    /* ---- tables to shared memory ---- */
    for (int i = tid; i < MP3T_HUFF_LUT_SIZE; i += blockDim.x) T.lut[i] = MP3T_HUFF_LUT[i];
    for (int i = tid; i < 145; i += blockDim.x) T.pow43[i] = MP3T_POW43[i];
    if (tid < 32) T.tabinfo[tid] = (uint32_t)MP3T_HUFF_BASE[tid] | ((uint32_t)MP3T_HUFF_ROOTW[tid] << 16) | ((uint32_t)MP3T_HUFF_LINBITS[tid] << 20);
    if (tid < 64) T.count1a[tid] = MP3T_COUNT1A[tid];
    for (int i = tid; i < 8*23; i += blockDim.x) T.sfb_long[i] = MP3T_SFB_LONG[i];
    for (int i = tid; i < 8*40; i += blockDim.x) { T.sfb_short[i] = MP3T_SFB_SHORT[i]; T.sfb_mixed[i] = MP3T_SFB_MIXED[i]; }
    __syncthreads();

    WarpScratch &W = scratch[warp];
    const uint32_t n_groups = (a.n_gc + 31)/32;

    for (uint32_t group = blockIdx.x*kWarpsPerBlock + warp; group < n_groups; group += gridDim.x*kWarpsPerBlock)
    {
        const uint32_t gc = group*32 + lane;
        const bool valid = gc < a.n_gc;
        GcDesc d;
        if (valid)
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(a.gcs + gc);
            uint4 *dst = reinterpret_cast<uint4 *>(&d);
            dst[0] = __ldg(src);
            dst[1] = __ldg(src + 1);
        } else
        {
            uint4 z = make_uint4(0, 0, 0, 0);
            reinterpret_cast<uint4 *>(&d)[0] = z;
            reinterpret_cast<uint4 *>(&d)[1] = z;
            d.flags1 = GCF1_MPEG1;
        }
        const bool mpeg1 = d.flags1 & GCF1_MPEG1;
        const int n_sfb = d.n_long_sfb + d.n_short_sfb;

        BitRd br;
        /* ---- scalefactors ---- */
        {
            const bool want_ist = valid && a.ist_pos && (((d.flags0 & GCF_I_STEREO) && (d.flags1 & GCF1_CH)) || a.tap_scf);
            uint8_t ist_local[40];
            const uint8_t *prev = nullptr;
            if (valid && (d.flags1 & GCF1_GR1) && d.scfsi)
            {
                /* scfsi: bands shared with granule 0 of the same channel -> re-read that granule's part 2 */
                const int nch = (d.flags1 & GCF1_STEREO) ? 2 : 1;
                GcDesc d0;
                const uint4 *src = reinterpret_cast<const uint4 *>(a.gcs + (gc - nch));
                reinterpret_cast<uint4 *>(&d0)[0] = __ldg(src);
                reinterpret_cast<uint4 *>(&d0)[1] = __ldg(src + 1);
                const ScfLayout L0 = scf_layout(d0);
                br.init(a.md_words, a.md_n_words, d0.bit_off);
                read_scalefactors(br, L0, d0.scfsi, false, W.prev, nullptr, nullptr, lane);
                prev = W.prev;
            }
            const ScfLayout L = scf_layout(d);
            br.init(a.md_words, a.md_n_words, d.bit_off);
            const int n_read = read_scalefactors(br, L, mpeg1 ? d.scfsi : 0, !mpeg1, W.iscf, prev, want_ist ? ist_local : nullptr, lane);
            if (want_ist)
            {
                uint8_t *o = a.ist_pos + (size_t)gc*40;
                for (int k = 0; k < 40; k++) o[k] = k < n_read ? ist_local[k] : (uint8_t)0;
            }
            /* subblock gain / preemphasis (minimp3.h:690-706), uint8 arithmetic wraps like the reference's */
            const int scf_shift = ((d.flags0 & GCF_SCALEFAC_SCALE) ? 1 : 0) + 1;
            if (d.n_short_sfb)
            {
                const int sh = 3 - scf_shift;
                for (int i = 0; i < d.n_short_sfb; i += 3)
                {
                    W.iscf[(d.n_long_sfb + i + 0)*32 + lane] += (uint8_t)(d.subblock_gain[0] << sh);
                    W.iscf[(d.n_long_sfb + i + 1)*32 + lane] += (uint8_t)(d.subblock_gain[1] << sh);
                    W.iscf[(d.n_long_sfb + i + 2)*32 + lane] += (uint8_t)(d.subblock_gain[2] << sh);
                }
            } else if (d.flags0 & GCF_PREFLAG)
            {
                for (int i = 0; i < 10; i++) W.iscf[(11 + i)*32 + lane] += c_preamp[i];
            }
        }
        const int scf_shift = ((d.flags0 & GCF_SCALEFAC_SCALE) ? 1 : 0) + 1;
        /* global gain (minimp3.h:708-709): BITS_DEQUANTIZER_OUT = -1, MAX_SCFI = 44 */
        const int gain_exp = (int)d.global_gain - 4 - 210 - ((d.flags0 & GCF_MS_STEREO) ? 2 : 0);
        const float gain = l3_ldexp_q2(2048.0f, 44 - gain_exp);
        if (a.tap_scf && valid)
        {
            float *o = a.tap_scf + (size_t)gc*40;
            for (int i = 0; i < 40; i++) o[i] = i < n_sfb ? l3_ldexp_q2(gain, (int)W.iscf[i*32 + lane] << scf_shift) : 0.f;
        }

        /* ---- Huffman + requantise ---- */
        const uint8_t *sfbw = d.n_short_sfb ? (d.n_long_sfb ? T.sfb_mixed + d.sr_idx*40 : T.sfb_short + d.sr_idx*40)
                                            : T.sfb_long + d.sr_idx*23;
        const int limit = d.part_23_length;
        int mode = d.big_values ? MODE_BIG : MODE_C1A;
        if (!valid) mode = MODE_DONE;
        int big_left = d.big_values;
        int band_left = 0, band_idx = 0;
        int ireg = 0, reg_left = d.region_count[0] + 1;
        uint32_t tinfo = T.tabinfo[d.table_select[0]];
        float one = 0.f;
        int pending = 0;
        int extent = 0;
        int16_t *tq = (a.tap_q && valid) ? a.tap_q + (size_t)gc*576 : nullptr;

        int iter = 0;
        int flushed = 0;   /* lines already flushed */
        for (;;)
        {
            float v0 = 0.f, v1 = 0.f;
            int q0 = 0, q1 = 0;
            if (mode != MODE_DONE)
            {
                bool run = true;
                if (mode == MODE_C1A)
                {
                    /* count1 quad code; stop when it ends past the granule's bit budget (minimp3.h:852-864) */
                    br.refill();
                    int len, vwxy;
                    if (d.flags0 & GCF_COUNT1_TABLE) { len = 4; vwxy = (~br.peek(4)) & 15; }
                    else { const int e = T.count1a[br.peek(6)]; len = e >> 4; vwxy = e & 15; }
                    br.skip(len);
                    if (br.pos > limit) { run = false; mode = MODE_DONE; }
                    pending = vwxy;
                }
                if (run && band_left == 0)
                {
                    const int wd = sfbw[band_idx];
                    if (!wd) { run = false; mode = MODE_DONE; }
                    else
                    {
                        band_left = wd >> 1;
                        one = l3_ldexp_q2(gain, (int)W.iscf[band_idx*32 + lane] << scf_shift);
                        band_idx++;
                        if (mode == MODE_BIG)
                        {
                            if (reg_left == 0)
                            {
                                ireg++;
                                tinfo = T.tabinfo[d.table_select[ireg]];
                                reg_left = d.region_count[ireg] + 1;
                            }
                            reg_left--;
                        }
                    }
                }
                if (run)
                {
                    if (mode == MODE_BIG)
                    {
                        const int base = tinfo & 0xFFFF, rootw = (tinfo >> 16) & 15, linbits = (tinfo >> 20) & 15;
                        int x = 0, y = 0;
                        br.refill();
                        if (rootw)
                        {
                            int e = T.lut[base + br.peek(rootw)];
                            int wcur = rootw;
                            while (e & 0x8000)
                            {
                                br.skip(wcur);
                                wcur = (e >> 11) & 15;
                                e = T.lut[base + (e & 0x7FF) + br.peek(wcur)];
                            }
                            br.skip(e >> 8);
                            x = (e >> 4) & 15;
                            y = e & 15;
                        }
                        br.refill();
                        if (x == 15 && linbits) { x += br.peek(linbits); br.skip(linbits); br.refill(); }
                        if (x)
                        {
                            const int neg = br.peek(1); br.skip(1);
                            const float m = MP3_FMUL(one, l3_pow43(x, T.pow43));
                            v0 = neg ? -m : m; q0 = neg ? -x : x;
                        }
                        br.refill();
                        if (y == 15 && linbits) { y += br.peek(linbits); br.skip(linbits); br.refill(); }
                        if (y)
                        {
                            const int neg = br.peek(1); br.skip(1);
                            const float m = MP3_FMUL(one, l3_pow43(y, T.pow43));
                            v1 = neg ? -m : m; q1 = neg ? -y : y;
                        }
                        if (--big_left == 0) mode = MODE_C1A;
                    } else
                    {
                        const int hi = (mode == MODE_C1A) ? 8 : 2;
                        br.refill();
                        if (pending & hi) { const int neg = br.peek(1); br.skip(1); v0 = neg ? -one : one; q0 = neg ? -1 : 1; }
                        if (pending & (hi >> 1)) { const int neg = br.peek(1); br.skip(1); v1 = neg ? -one : one; q1 = neg ? -1 : 1; }
                        mode = (mode == MODE_C1A) ? MODE_C1B : MODE_C1A;
                    }
                    band_left--;
                    extent = 2*iter + 2;
                    if (tq) { tq[2*iter] = (int16_t)q0; tq[2*iter + 1] = (int16_t)q1; }
                }
            }
            {
                const int col = (2*iter) & (kStageLines - 1);
                W.stage[col*kStagePitch + lane] = v0;
                W.stage[(col + 1)*kStagePitch + lane] = v1;
            }
            iter++;
            const bool boundary16 = (iter & 15) == 0;
            bool finished = iter == 288;
            if (boundary16 && !finished) finished = __ballot_sync(0xffffffffu, mode != MODE_DONE) == 0;
            if (finished || (iter & 31) == 0)
            {
                /* ---- flush lines [flushed, 2*iter) of all 32 rows ---- */
                __syncwarp();
                const int my_ext = (mode == MODE_DONE) ? ((extent + 31) & ~31) : 576;
                const int my_ms = (d.flags1 & GCF1_MS_PLAIN) ? 1 : 0;
                const int n_lines = 2*iter - flushed;   /* 32 or 64 */
                for (int r = 0; r < 32; r += 2)
                {
                    int e0 = __shfl_sync(0xffffffffu, my_ext, r);
                    int e1 = __shfl_sync(0xffffffffu, my_ext, r + 1);
                    const int ms = __shfl_sync(0xffffffffu, my_ms, r);
                    if (ms) e0 = e1 = max(e0, e1);
                    if (flushed >= max(e0, e1)) continue;
                    const size_t row0 = (size_t)(group*32 + r)*576;
                    for (int c = lane; c < n_lines; c += 32)
                    {
                        const int line = flushed + c;
                        const int col = line & (kStageLines - 1);
                        float l = W.stage[col*kStagePitch + r];
                        float rr = W.stage[col*kStagePitch + r + 1];
                        if (ms) { const float s = l + rr, dd = l - rr; l = s; rr = dd; }
                        if (line < e0) a.xr[row0 + line] = l;
                        if (line < e1) a.xr[row0 + 576 + line] = rr;
                    }
                }
                flushed = 2*iter;
                __syncwarp();
            }
            if (finished) break;
        }
        /* ---- per-row extent for the consumer ---- */
        {
            int e = (extent + 31) & ~31;
            const int other = __shfl_xor_sync(0xffffffffu, e, 1);
            if (d.flags1 & GCF1_MS_PLAIN) e = max(e, other);
            if (valid) a.nzl[gc] = (uint16_t)e;
        }
        __syncwarp();
    }
}

}  // namespace

cudaError_t launch_entropy(const DeviceTables &, const EntropyArgs &a, cudaStream_t s)
{
    if (!a.n_gc) return cudaSuccess;
    const size_t smem = ((sizeof(SharedTables) + 15) & ~15) + sizeof(WarpScratch)*kWarpsPerBlock;
    cudaError_t e = cudaFuncSetAttribute(k_l3_entropy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t n_groups = (a.n_gc + 31)/32;
    uint32_t blocks = (n_groups + kWarpsPerBlock - 1)/kWarpsPerBlock;
    const uint32_t max_blocks = (uint32_t)sms*2;
    if (blocks > max_blocks) blocks = max_blocks;
    k_l3_entropy<<<blocks, kWarpsPerBlock*32, smem, s>>>(a);
    return cudaGetLastError();
}

}  // namespace mp3b
