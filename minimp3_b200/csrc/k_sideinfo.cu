/*
 * k_sideinfo.cu -- Layer III side info -> GcDesc records on the GPU.
 *
 * The host walker only does what is sequential -- frame sync, the reservoir accounting, tags -- and reads a third
 * of the side-info fields for it (read_side_info_lite_t).  It hands every decodable frame over as a 64-byte
 * FrameRec with the raw side-info bytes; one thread per frame runs the full parse of L3_read_side_info
 * (minimp3.h:484-607, restated in l3_sideinfo.cuh) and writes the 2-4 descriptors of the frame.
 */
#include <cuda_runtime.h>

#include "kernels.h"
#include "l3_sideinfo.cuh"

namespace mp3b {

namespace {

/* the 32 side-info bytes as four 64-bit words in bitstream order: the parse reads its ~60 fields from registers (the
 * byte-wise reader of the host build turns into local-memory traffic here: 1500 LDL per thread) */
struct SiRegBits
{
    unsigned long long w[4];
    int pos, limit;
    __device__ __forceinline__ unsigned long long word(int i) const
    {
        return i == 0 ? w[0] : i == 1 ? w[1] : i == 2 ? w[2] : i == 3 ? w[3] : 0ull;
    }
    __device__ __forceinline__ uint32_t get(int n)
    {
        const int s = pos;
        pos += n;
        if (pos > limit || n == 0) return 0;
        const int i = s >> 6, sh = s & 63;
        const unsigned long long hi = word(i), lo = word(i + 1);
        const unsigned long long v = sh ? (hi << sh) | (lo >> (64 - sh)) : hi;
        return (uint32_t)(v >> (64 - n));
    }
};

/* frame_descs (l3_sideinfo.cuh) over the register reader */
__device__ __forceinline__ int frame_descs_reg(const FrameRec &fr, GcDesc *out)
{
    SiRegBits bs;
    const uint32_t *si = reinterpret_cast<const uint32_t *>(fr.si);
#pragma unroll
    for (int k = 0; k < 4; k++)
        bs.w[k] = (unsigned long long)__byte_perm(si[2*k], 0, 0x0123) << 32 | __byte_perm(si[2*k + 1], 0, 0x0123);
    bs.pos = 0;
    bs.limit = 256;
    GranuleSide side[4];
    if (read_side_info_t<false>(&bs, side, fr.hdr) < 0) return 0;
    const int nch = fr.nch, n_gr = fr.n_gr;
    const HdrBits hb(fr.hdr, nch);
    uint64_t bit = fr.md_bit;
    for (int igr = 0; igr < n_gr; igr++)
        for (int ch = 0; ch < nch; ch++)
        {
            const GranuleSide &g = side[igr*nch + ch];
            fill_gc_desc_hb(&out[igr*nch + ch], g, hb, ch, igr, bit);
            bit += g.part_23_length;
        }
    return n_gr*nch;
}

__global__ void __launch_bounds__(256) k_l3_side_info(const FrameRec *__restrict__ frames, uint32_t n_frames, GcDesc *__restrict__ gcs)
{
    const uint32_t f = blockIdx.x*blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    FrameRec fr;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(frames + f);
        uint4 *dst = reinterpret_cast<uint4 *>(&fr);
#pragma unroll
        for (int k = 0; k < 4; k++) dst[k] = __ldg(src + k);
    }
    GcDesc d[4];
    const int n = frame_descs_reg(fr, d);
    uint4 *out = reinterpret_cast<uint4 *>(gcs + fr.gc_first);
    /* descriptors of granule 0 and granule 1 are contiguous: gc_first .. gc_first + n_gr*nch - 1 */
    for (int k = 0; k < n; k++)
    {
        out[2*k] = reinterpret_cast<const uint4 *>(&d[k])[0];
        out[2*k + 1] = reinterpret_cast<const uint4 *>(&d[k])[1];
    }
    if (fr.pad[0])
    {
        /* the walker skipped one slot in front of this frame to keep stereo pairs on even indices */
        GcDesc pad;
        memset(&pad, 0, sizeof(pad));
        pad.flags1 = GCF1_MPEG1;
        uint4 *po = reinterpret_cast<uint4 *>(gcs + fr.gc_first - 1);
        po[0] = reinterpret_cast<const uint4 *>(&pad)[0];
        po[1] = reinterpret_cast<const uint4 *>(&pad)[1];
    }
}

}  // namespace

cudaError_t launch_side_info(const FrameRec *frames, uint32_t n_frames, GcDesc *gcs, cudaStream_t s)
{
    if (!n_frames) return cudaSuccess;
    k_l3_side_info<<<(n_frames + 255)/256, 256, 0, s>>>(frames, n_frames, gcs);
    return cudaGetLastError();
}

}  // namespace mp3b
