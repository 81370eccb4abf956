/*
 * k_stereo.cu -- intensity stereo (and the MS bands that live next to it) for the granules whose
 * frame header has the intensity bit set.  Restates L3_intensity_stereo, L3_stereo_top_band,
 * L3_stereo_process and L3_intensity_stereo_band (minimp3.h:911-993).
 *
 * Only joint-stereo streams that actually use intensity coding come here: one warp per granule, lanes over lines.
 * The two rows are staged in shared memory with coalesced loads; while they pass, every lane notes the bands of the
 * right channel's non-zero lines in a 64-bit mask (line -> band through a table per sample rate and block kind), so
 * the reference's band-by-band scan for the top band becomes one OR across the warp; the per-band decisions
 * (pass through / MS / intensity, the two gains) are taken by one lane per band, and the lines are rewritten in one
 * sweep.  (Walking the bands one after the other with the lanes inside a band -- most bands are narrower than a
 * warp -- cost 3200 warp instructions per granule and made this kernel issue-bound: profiles/r02_stereo_cfg5.md.)
 */
#include <cuda_runtime.h>

#include "kernels.h"
#include "l3_math.cuh"

namespace mp3b {

namespace {

constexpr int kWarps = 4;

/* ISO 11172-3 2.4.3.4.9.3: is_ratio = tan(is_pos*pi/12); left = r/(1+r), right = 1/(1+r) */
__constant__ float c_pan[14] = { 0.f, 1.f, 0.21132487f, 0.78867513f, 0.36602540f, 0.63397460f, 0.5f, 0.5f,
                                 0.63397460f, 0.36602540f, 0.78867513f, 0.21132487f, 1.f, 0.f };

struct WarpShared
{
    float4 rows[2][144];
    float kl[40], kr[40];
    uint8_t mode[40];       /* 0 pass through, 1 MS, 2 intensity */
    uint8_t ist[40];
};

__global__ void __launch_bounds__(kWarps*32) k_l3_intensity_stereo(StereoArgs a, const uint8_t *__restrict__ band_map)
{
    __shared__ WarpShared sh[kWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t gi = blockIdx.x*kWarps + warp;
    if (gi >= a.n_granules) return;
    WarpShared &S = sh[warp];
    const uint32_t gc0 = a.granules[gi], gc1 = gc0 + 1;
    const GcDesc d0 = a.gcs[gc0], d1 = a.gcs[gc1];
    const bool mpeg1 = d0.flags1 & GCF1_MPEG1;
    const bool ms_bit = d0.flags1 & GCF1_MS_TEST;
    const int n_long = d0.n_long_sfb, n_short = d0.n_short_sfb, n_sfb = n_long + n_short;
    const int max_blocks = n_short ? 3 : 1;
    const int kind = n_short ? (n_long ? 1 : 2) : 0;
    const uint8_t *bmap = band_map + (kind*8 + d0.sr_idx)*576;
    float *L = a.xr + (size_t)gc0*576, *R = a.xr + (size_t)gc1*576;
    const int nzl0 = a.nzl[gc0], nzl1 = a.nzl[gc1];
    const int ext = max(nzl0, nzl1);

    for (int k = lane; k < 40; k += 32) S.ist[k] = a.ist_pos[(size_t)gc1*40 + k];
    /* L3_stereo_top_band: last band (per window class) of the right channel holding a non-zero line */
    unsigned long long nz = 0ull;
    /* extents are multiples of 32 lines (k_entropy.cu): a lane moves four lines at a time */
    const int ext4 = ext >> 2, nzl0_4 = nzl0 >> 2, nzl1_4 = nzl1 >> 2;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 *L4 = reinterpret_cast<const float4 *>(L), *R4 = reinterpret_cast<const float4 *>(R);
    const uint32_t *bmap4 = reinterpret_cast<const uint32_t *>(bmap);
#pragma unroll 2
    for (int k = lane; k < ext4; k += 32)
    {
        const float4 l = k < nzl0_4 ? L4[k] : zero4;     /* lines at or beyond a row's extent were never written: zero */
        const float4 r = k < nzl1_4 ? R4[k] : zero4;
        S.rows[0][k] = l;
        S.rows[1][k] = r;
        const uint32_t bb = __ldg(bmap4 + k);
        if (r.x != 0.f) nz |= 1ull << (bb & 255u);
        if (r.y != 0.f) nz |= 1ull << ((bb >> 8) & 255u);
        if (r.z != 0.f) nz |= 1ull << ((bb >> 16) & 255u);
        if (r.w != 0.f) nz |= 1ull << (bb >> 24);
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) nz |= __shfl_xor_sync(0xffffffffu, nz, off);
    if (n_sfb < 64) nz &= (1ull << n_sfb) - 1;
    int max_band[3];
#pragma unroll
    for (int c = 0; c < 3; c++)
    {
        const unsigned long long m = nz & (0x9249249249249249ull << c);      /* bands i with i % 3 == c */
        max_band[c] = m ? 63 - __clzll((long long)m) : -1;
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
            S.ist[itop] = (uint8_t)(max_band[i] >= prev ? default_pos : S.ist[prev]);
        }
    }
    __syncwarp();

    /* L3_stereo_process: the decision and the gains of every band, one lane per band */
    const unsigned max_pos = mpeg1 ? 7 : 64;
    const int mpeg2_sh = d1.scalefac_compress & 1;
    for (int i = lane; i < 40; i += 32)
    {
        const unsigned ipos = S.ist[i];
        const int top = i % 3 == 0 ? max_band[0] : i % 3 == 1 ? max_band[1] : max_band[2];
        const bool intensity = i > top && ipos < max_pos;
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
        S.kl[i] = kl;
        S.kr[i] = kr;
        S.mode[i] = (uint8_t)(intensity ? 2 : ms_bit ? 1 : 0);
    }
    __syncwarp();
    auto line = [&](int b, float l, float r, float &nl, float &nr) {
        const int mode = S.mode[b];
        nl = l; nr = r;
        if (mode == 2) { nr = MP3_FMUL(l, S.kr[b]); nl = MP3_FMUL(l, S.kl[b]); }
        else if (mode == 1) { nl = MP3_FADD(l, r); nr = MP3_FADD(l, -r); }
    };
#pragma unroll 2
    for (int k = lane; k < ext4; k += 32)
    {
        const uint32_t bb = __ldg(bmap4 + k);
        const float4 l = S.rows[0][k], r = S.rows[1][k];
        float4 nl, nr;
        line((int)(bb & 255u), l.x, r.x, nl.x, nr.x);
        line((int)((bb >> 8) & 255u), l.y, r.y, nl.y, nr.y);
        line((int)((bb >> 16) & 255u), l.z, r.z, nl.z, nr.z);
        line((int)(bb >> 24), l.w, r.w, nl.w, nr.w);
        reinterpret_cast<float4 *>(L)[k] = nl;
        reinterpret_cast<float4 *>(R)[k] = nr;
    }
    if (lane == 0)
    {
        a.nzl[gc0] = (uint16_t)ext;
        a.nzl[gc1] = (uint16_t)ext;
    }
}

}  // namespace

cudaError_t launch_intensity_stereo(const DeviceTables &t, const StereoArgs &a, cudaStream_t s)
{
    if (!a.n_granules) return cudaSuccess;
    const uint32_t blocks = (a.n_granules + kWarps - 1)/kWarps;
    k_l3_intensity_stereo<<<blocks, kWarps*32, 0, s>>>(a, t.band_of_line);
    return cudaGetLastError();
}

}  // namespace mp3b
