/*
 * k_stereo.cu -- intensity stereo (and the MS bands that live next to it) for the granules whose
 * frame header has the intensity bit set.  Restates L3_intensity_stereo, L3_stereo_top_band,
 * L3_stereo_process and L3_intensity_stereo_band (minimp3.h:911-993).
 *
 * Only joint-stereo streams that actually use intensity coding come here: one warp per granule.  The two rows
 * are staged in shared memory with coalesced loads first -- the band walk (up to 39 bands, twice) is a chain of
 * short dependent steps and must not pay a global round trip per band -- and written back in one sweep.
 */
#include <cuda_runtime.h>

#include "kernels.h"
#include "l3_math.cuh"
#define MP3T_QUAL static __device__ const
#include "mp3_tables.h"

namespace mp3b {

namespace {

constexpr int kWarps = 4;

/* ISO 11172-3 2.4.3.4.9.3: is_ratio = tan(is_pos*pi/12); left = r/(1+r), right = 1/(1+r) */
__constant__ float c_pan[14] = { 0.f, 1.f, 0.21132487f, 0.78867513f, 0.36602540f, 0.63397460f, 0.5f, 0.5f,
                                 0.63397460f, 0.36602540f, 0.78867513f, 0.21132487f, 1.f, 0.f };

__global__ void __launch_bounds__(kWarps*32) k_l3_intensity_stereo(StereoArgs a)
{
    __shared__ uint8_t s_ist[kWarps][40];
    __shared__ float s_rows[kWarps][2][576];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t gi = blockIdx.x*kWarps + warp;
    if (gi >= a.n_granules) return;
    const uint32_t gc0 = a.granules[gi], gc1 = gc0 + 1;
    const GcDesc d0 = a.gcs[gc0], d1 = a.gcs[gc1];
    const bool mpeg1 = d0.flags1 & GCF1_MPEG1;
    const bool ms_bit = d0.flags1 & GCF1_MS_TEST;
    const int n_long = d0.n_long_sfb, n_short = d0.n_short_sfb, n_sfb = n_long + n_short;
    const int max_blocks = n_short ? 3 : 1;
    const uint8_t *sfbw = n_short ? (n_long ? MP3T_SFB_MIXED + d0.sr_idx*40 : MP3T_SFB_SHORT + d0.sr_idx*40)
                                  : MP3T_SFB_LONG + d0.sr_idx*23;
    float *L = a.xr + (size_t)gc0*576, *R = a.xr + (size_t)gc1*576;
    const int nzl0 = a.nzl[gc0], nzl1 = a.nzl[gc1];
    const int ext = max(nzl0, nzl1);

    for (int k = lane; k < 40; k += 32) s_ist[warp][k] = a.ist_pos[(size_t)gc1*40 + k];
    float *Ls = s_rows[warp][0], *Rs = s_rows[warp][1];
    for (int k = lane; k < ext; k += 32)
    {
        Ls[k] = k < nzl0 ? L[k] : 0.f;       /* lines at or beyond a row's extent were never written: zero */
        Rs[k] = k < nzl1 ? R[k] : 0.f;
    }
    __syncwarp();

    /* L3_stereo_top_band: last band (per window class) of the right channel holding a non-zero line */
    int max_band[3] = { -1, -1, -1 };
    {
        int start = 0;
        for (int i = 0; i < n_sfb && start < nzl1; i++)
        {
            const int w = sfbw[i];
            bool nz = false;
            for (int k = lane; k < w; k += 32)
            {
                const int line = start + k;
                if (line < nzl1 && Rs[line] != 0.f) nz = true;
            }
            if (__any_sync(0xffffffffu, nz)) max_band[i % 3] = i;
            start += w;
        }
    }
    if (n_long)
    {
        const int m = max(max(max_band[0], max_band[1]), max_band[2]);
        max_band[0] = max_band[1] = max_band[2] = m;
    }
    if (lane == 0)
    {
        for (int i = 0; i < max_blocks; i++)
        {
            const int default_pos = mpeg1 ? 3 : 0;
            const int itop = n_sfb - max_blocks + i;
            const int prev = itop - max_blocks;
            s_ist[warp][itop] = (uint8_t)(max_band[i] >= prev ? default_pos : s_ist[warp][prev]);
        }
    }
    __syncwarp();

    /* L3_stereo_process */
    const unsigned max_pos = mpeg1 ? 7 : 64;
    const int mpeg2_sh = d1.scalefac_compress & 1;
    int start = 0;
    for (int i = 0; sfbw[i] && start < ext; i++)
    {
        const int w = sfbw[i];
        const unsigned ipos = s_ist[warp][i];
        const bool intensity = i > max_band[i % 3] && ipos < max_pos;
        float kl = 0.f, kr = 0.f;
        if (intensity)
        {
            const float s = ms_bit ? 1.41421356f : 1.f;
            if (mpeg1)
            {
                kl = c_pan[2*ipos];
                kr = c_pan[2*ipos + 1];
            } else
            {
                kl = 1.f;
                kr = l3_ldexp_q2(1.f, (int)((ipos + 1) >> 1 << mpeg2_sh));
                if (ipos & 1) { kl = kr; kr = 1.f; }
            }
            kl = MP3_FMUL(kl, s);
            kr = MP3_FMUL(kr, s);
        }
        for (int k = lane; k < w; k += 32)
        {
            const int line = start + k;
            if (line >= ext) continue;
            const float l = Ls[line], r = Rs[line];
            float nl = l, nr = r;
            if (intensity) { nr = MP3_FMUL(l, kr); nl = MP3_FMUL(l, kl); }
            else if (ms_bit) { nl = MP3_FADD(l, r); nr = MP3_FADD(l, -r); }
            Ls[line] = nl;
            Rs[line] = nr;
        }
        start += w;
    }
    __syncwarp();
    for (int k = lane; k < ext; k += 32)
    {
        L[k] = Ls[k];
        R[k] = Rs[k];
    }
    if (lane == 0)
    {
        a.nzl[gc0] = (uint16_t)ext;
        a.nzl[gc1] = (uint16_t)ext;
    }
}

}  // namespace

cudaError_t launch_intensity_stereo(const DeviceTables &, const StereoArgs &a, cudaStream_t s)
{
    if (!a.n_granules) return cudaSuccess;
    const uint32_t blocks = (a.n_granules + kWarps - 1)/kWarps;
    k_l3_intensity_stereo<<<blocks, kWarps*32, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace mp3b
