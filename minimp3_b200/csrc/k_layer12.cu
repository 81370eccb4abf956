/*
 * k_layer12.cu -- Layer I / II frame payload decode: bit allocation + scalefactors + sample unpacking,
 * scaled and written as subband samples sub12[block-channel][12 slots][32 subbands] for the synthesis kernel.
 * Restates L12_read_scale_info, L12_dequantize_granule and L12_apply_scf_384 (minimp3.h:362-481); the
 * arithmetic lives in l12_decode.cuh.
 *
 * Mapping: a block of 384 threads takes 32 frames.  One lane per frame parses the scale info (serial, a few
 * hundred bits) into shared memory; then 12 threads per frame each decode one sub-block of samples -- the
 * sub-blocks of a frame all have the same bit length, so their bit offsets are known without decoding.
 */
#include <cuda_runtime.h>

#include "kernels.h"
#include "l12_decode.cuh"
#include "l3_entropy.cuh"

namespace mp3b {

namespace {

constexpr int kFramesPerBlock = 32;

struct FrameShared
{
    L12Info sci;
    uint64_t sample_bit_off;     /* absolute bit offset of the first sub-block */
    int sub_bits;
    int group;
};

__global__ void __launch_bounds__(kFramesPerBlock*12) k_l12_dequant(Layer12Args a)
{
    __shared__ FrameShared sh[kFramesPerBlock];
    const int tid = threadIdx.x;
    const uint32_t frame0 = blockIdx.x*kFramesPerBlock;

    if (tid < kFramesPerBlock && frame0 + tid < a.n_frames)
    {
        const L12FrameDesc d = a.frames[frame0 + tid];
        BitRd br;
        br.init(a.md_words, a.md_n_words, d.bit_off);
        FrameShared &F = sh[tid];
        l12_read_scale_info(d.hdr, d.kbps, br, &F.sci, true);
        F.group = d.n_blocks == 1 ? 1 : 3;
        F.sub_bits = l12_subblock_bits(F.sci, F.group);
        F.sample_bit_off = d.bit_off + (uint64_t)br.bits_consumed();
    }
    __syncthreads();

    const int f = tid/12, k = tid - f*12;
    if (frame0 + f >= a.n_frames) return;
    const L12FrameDesc d = a.frames[frame0 + f];
    const FrameShared &F = sh[f];
    BitRd br;
    br.init(a.md_words, a.md_n_words, F.sample_bit_off + (uint64_t)k*(uint64_t)F.sub_bits);
    /* Layer II: sub-block k = part k/4 (own scalefactor), slots 3(k%4)..; Layer I: slot k of the only block */
    const int part = F.group == 3 ? k >> 2 : 0;
    const int igr = k >> 2;
    const int slot0 = F.group == 3 ? 3*(k & 3) : k;
    const int nch = d.nch;
    float *base = a.sub12 + (size_t)(d.block_first + part*nch)*384;
    l12_dequant_subblock(br, F.sci, F.group, igr, slot0, [&](int ch, int slot, int sb, float v) {
        if (ch < nch) base[(size_t)ch*384 + slot*32 + sb] = v;
    });
}

}  // namespace

cudaError_t launch_l12_dequant(const Layer12Args &a, cudaStream_t s)
{
    if (!a.n_frames) return cudaSuccess;
    const uint32_t blocks = (a.n_frames + kFramesPerBlock - 1)/kFramesPerBlock;
    k_l12_dequant<<<blocks, kFramesPerBlock*12, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace mp3b
