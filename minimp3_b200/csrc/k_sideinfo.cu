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
    const int n = frame_descs(fr, d);
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
