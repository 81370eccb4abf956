/*
 * k_layer12.cu -- Layer I / II frame payload decode: bit allocation + scalefactors + sample unpacking,
 * scaled and written as subband samples sub12[block-channel][12 slots][32 subbands] for the synthesis kernel.
 * Restates L12_read_scale_info, L12_dequantize_granule and L12_apply_scf_384 (minimp3.h:362-481); the
 * arithmetic lives in l12_decode.cuh.
 *
 * Mapping: two kernels.  k_l12_scale_info: one thread per frame parses the scale info (serial, a few hundred
 * bits, latency-bound: it wants every lane of many warps busy) and leaves it in compact form in HBM (272 bytes per
 * frame).  k_l12_dequant: a block of 384 threads takes 32 frames, expands their scale info in shared memory (the
 * scalefactor values are computed here, in parallel), then 12 threads per frame each decode one sub-block of
 * samples -- the sub-blocks of a frame all have the same bit length, so their bit offsets are known without
 * decoding.  (One kernel doing both had 11 of 12 warps waiting at the barrier behind the parse: issue slots 14 %
 * busy, profiles/r02_l12_dequant.md.)
 */
#include <cuda_runtime.h>

#include "kernels.h"
#include "l12_decode.cuh"
#include "l3_entropy.cuh"

namespace mp3b {

namespace {

constexpr int kFramesPerBlock = 32;
constexpr int kScaleThreads = 128;

__global__ void __launch_bounds__(kScaleThreads) k_l12_scale_info(Layer12Args a)
{
    const uint32_t f = blockIdx.x*kScaleThreads + threadIdx.x;
    if (f >= a.n_frames) return;
    const L12FrameDesc d = a.frames[f];
    BitRd br;
    br.init(a.md_words, a.md_n_words, d.bit_off);
    L12Compact c;
    {
        uint4 *z = reinterpret_cast<uint4 *>(&c);
        for (int i = 0; i < 4; i++) z[i] = make_uint4(0u, 0u, 0u, 0u);
        for (int i = 4; i < 16; i++) z[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
    }
    l12_parse_scale_info(d.hdr, d.kbps, br, &c, [&](int idx, int index, int) { c.scf_index[idx] = (uint8_t)(index < 0 ? 0xFF : index); });
    c.group = (uint8_t)(d.n_blocks == 1 ? 1 : 3);
    c.pad = 0;
    c.sub_bits = l12_subblock_bits(c, c.group);
    c.sample_bit_off = d.bit_off + (uint64_t)br.bits_consumed();
    uint4 *o = reinterpret_cast<uint4 *>(a.info + f);
    const uint4 *src = reinterpret_cast<const uint4 *>(&c);
    for (int i = 0; i < 17; i++) o[i] = src[i];
}

__global__ void __launch_bounds__(kFramesPerBlock*12) k_l12_dequant(Layer12Args a)
{
    __shared__ L12Info sh[kFramesPerBlock];
    __shared__ L12Compact shc[kFramesPerBlock];
    const int tid = threadIdx.x;
    const uint32_t frame0 = blockIdx.x*kFramesPerBlock;
    const int n_here = (int)min((uint32_t)kFramesPerBlock, a.n_frames - frame0);

    {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.info + frame0);
        uint4 *dst = reinterpret_cast<uint4 *>(shc);
        for (int i = tid; i < n_here*17; i += kFramesPerBlock*12) dst[i] = src[i];
    }
    __syncthreads();
    /* scalefactor values for every (frame, band-channel, part): 16 per thread */
    for (int i = tid; i < n_here*192; i += kFramesPerBlock*12)
    {
        const int f = i/192, j = i - f*192, bc = j/3;
        const L12Compact &c = shc[f];
        const int index = c.scf_index[j];
        sh[f].scf[j] = index == 0xFF ? 0.f : l12_scf_value(index, l12_scale_alloc(c.bitalloc, c.stereo_bands, bc));
    }
    for (int i = tid; i < n_here*64; i += kFramesPerBlock*12)
    {
        const int f = i >> 6, j = i & 63;
        sh[f].bitalloc[j] = shc[f].bitalloc[j];
        if (j == 0) { sh[f].total_bands = shc[f].total_bands; sh[f].stereo_bands = shc[f].stereo_bands; }
    }
    __syncthreads();

    const int f = tid/12, k = tid - f*12;
    if (f >= n_here) return;
    const L12FrameDesc d = a.frames[frame0 + f];
    const L12Info &sci = sh[f];
    const int group = shc[f].group;
    BitRd br;
    br.init(a.md_words, a.md_n_words, shc[f].sample_bit_off + (uint64_t)k*(uint64_t)shc[f].sub_bits);
    /* Layer II: sub-block k = part k/4 (own scalefactor), slots 3(k%4)..; Layer I: slot k of the only block */
    const int part = group == 3 ? k >> 2 : 0;
    const int igr = k >> 2;
    const int slot0 = group == 3 ? 3*(k & 3) : k;
    const int nch = d.nch;
    float *base = a.sub12 + (size_t)(d.block_first + part*nch)*384;
    /* four bands at a time: the samples leave as 16-byte stores (as single floats every store was a sector
     * transaction of its own at the L2 -- 470 M of them per launch, which is what the kernel's time was) */
    for (int b0 = 0; b0 < 32; b0 += 4)
    {
        float w[2][3][4];
#pragma unroll
        for (int q = 0; q < 4; q++)
        {
            float v[2][3];
            l12_dequant_band(br, sci, group, igr, b0 + q, v);
#pragma unroll
            for (int ch = 0; ch < 2; ch++)
#pragma unroll
                for (int kk = 0; kk < 3; kk++) w[ch][kk][q] = v[ch][kk];
        }
#pragma unroll
        for (int ch = 0; ch < 2; ch++)
#pragma unroll
            for (int kk = 0; kk < 3; kk++)
                if (ch < nch && kk < group)
                    *reinterpret_cast<float4 *>(base + (size_t)ch*384 + (slot0 + kk)*32 + b0) = make_float4(w[ch][kk][0], w[ch][kk][1], w[ch][kk][2], w[ch][kk][3]);
    }
}

}  // namespace

cudaError_t launch_l12_dequant(const Layer12Args &a, cudaStream_t s)
{
    if (!a.n_frames) return cudaSuccess;
    k_l12_scale_info<<<(a.n_frames + kScaleThreads - 1)/kScaleThreads, kScaleThreads, 0, s>>>(a);
    const uint32_t blocks = (a.n_frames + kFramesPerBlock - 1)/kFramesPerBlock;
    k_l12_dequant<<<blocks, kFramesPerBlock*12, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace mp3b
