/*
 * l12_decode.cuh -- Layer I / II frame payload decode (bit allocation, scalefactors, sample unpacking and
 * scaling) as __host__ __device__ code: the host uses the scale-info parser to know how many bits a frame
 * consumes (the reference drops frames that overrun), the CUDA lanes use all of it.
 *
 * Restates L12_subband_alloc_table, L12_read_scale_info, L12_read_scalefactors, L12_dequantize_granule and
 * L12_apply_scf_384 (minimp3.h:316-481).  The bit-allocation tables are ISO 11172-3 tables B.2a-d / 13818-3 B.1
 * written as numbers of quantisation levels per allocation code.
 */
#ifndef MP3B_L12_DECODE_CUH
#define MP3B_L12_DECODE_CUH
#include <stdint.h>

#include "l3_math.cuh"

namespace mp3b {

/* allocation rows: levels per 4-bit (or shorter) allocation code; 3, 5 and 9 levels are coded as grouped
 * triplets except in Layer I; row 6 is the Layer I table (code c -> c+1 bits) */
#define MP3B_L12_ROWS { \
    /* row 0 (offset  0) */ 0,3,7,15,31,63,127,255,511,1023,2047,4095,8191,16383,32767,65535, \
    /* row 1 (offset 16) */ 0,3,5,7,9,15,31,63,127,255,511,1023,2047,4095,8191,65535, \
    /* row 2 (offset 32) */ 0,3,5,7,9,15,31,65535, \
    /* row 3 (offset 40) */ 0,3,5,65535, \
    /* row 4 (offset 44) */ 0,3,5,9,15,31,63,127,255,511,1023,2047,4095,8191,16383,32767, \
    /* row 5 (offset 60) */ 0,3,5,7,9,15,31,63,127,255,511,1023,2047,4095,8191,16383 }
#if defined(__CUDACC__)
static __device__ const uint16_t d_l12_levels[76] = MP3B_L12_ROWS;
#endif
static const uint16_t h_l12_levels[76] = MP3B_L12_ROWS;

MP3_HD int l12_levels(int i)
{
#if defined(__CUDA_ARCH__)
    return d_l12_levels[i];
#else
    return h_l12_levels[i];
#endif
}

/* allocation code -> "ba": 0 none, 2..16 bits per sample, 17/18/19 grouped triplets of 3/5/9 levels */
MP3_HD int l12_ba(int row_off, int code, bool layer1)
{
    if (layer1) return code ? code + 1 : 0;
    const int lv = l12_levels(row_off + code);
    if (lv == 0) return 0;
    if (lv == 3) return 17;
    if (lv == 5) return 18;
    if (lv == 9) return 19;
    int bits = 2;
    while (((1 << bits) - 1) != lv) bits++;
    return bits;
}

struct L12Range { uint8_t row_off, width, count; };

struct L12Alloc
{
    L12Range r[4];
    int n_ranges, total_bands, stereo_bands;
    bool layer1;
};

/* L12_subband_alloc_table (minimp3.h:317-360) */
MP3_HD void l12_alloc_table(const uint8_t *hdr, unsigned kbps_total, L12Alloc *a)
{
    const int mode = (hdr[3] >> 6) & 3;
    const int stereo_bands = (mode == 3) ? 0 : (mode == 1) ? ((((hdr[3] >> 4) & 3) << 2) + 4) : 32;
    int nbands;
    a->layer1 = (hdr[1] & 6) == 6;
    if (a->layer1)
    {
        a->n_ranges = 1; a->r[0] = L12Range{ 0, 4, 32 };
        nbands = 32;
    } else if (!(hdr[1] & 8))
    {
        a->n_ranges = 3; a->r[0] = L12Range{ 60, 4, 4 }; a->r[1] = L12Range{ 44, 3, 7 }; a->r[2] = L12Range{ 44, 2, 19 };
        nbands = 30;
    } else
    {
        const int sr = (hdr[2] >> 2) & 3;
        unsigned kbps = kbps_total >> (mode != 3 ? 1 : 0);
        if (!kbps) kbps = 192;   /* free format */
        a->n_ranges = 4; a->r[0] = L12Range{ 0, 4, 3 }; a->r[1] = L12Range{ 16, 4, 8 }; a->r[2] = L12Range{ 32, 3, 12 }; a->r[3] = L12Range{ 40, 2, 7 };
        nbands = 27;
        if (kbps < 56)
        {
            a->n_ranges = 2; a->r[0] = L12Range{ 44, 4, 2 }; a->r[1] = L12Range{ 44, 3, 10 };
            nbands = sr == 2 ? 12 : 8;
        } else if (kbps >= 96 && sr != 1)
            nbands = 30;
    }
    a->total_bands = nbands;
    a->stereo_bands = stereo_bands < nbands ? stereo_bands : nbands;
}

struct L12Info
{
    uint8_t total_bands, stereo_bands;
    uint8_t bitalloc[64];
    float scf[3*64];
};

/* the scalefactor of 6-bit index b for a band quantised with allocation ba (minimp3.h:385-400):
 * 2^-20 * 2^(-(b%3)/3) / levels * 2^(21 - b/3) */
MP3_HD float l12_scf_value(int b, int ba)
{
    const int lv = ba < 17 ? (1 << ba) - 1 : (2 << (ba - 17)) + 1;
    const int r = b % 3;
    const float base = r == 0 ? 9.53674316e-07f : r == 1 ? 7.56931807e-07f : 6.00777173e-07f;
    return MP3_FMUL(MP3_FDIV(base, (float)lv), (float)((1 << 21) >> (b/3)));
}

/* L12_read_scale_info + L12_read_scalefactors (minimp3.h:362-432).  Reader needs get(n); Info needs total_bands,
 * stereo_bands and bitalloc[64]; scf_out(3*i + part, index, ba) receives the scalefactor index in force for every
 * (band-channel i, part) -- -1 where the band carries nothing -- and the allocation it scales. */
template <class Reader, class Info, class ScfOut>
MP3_HD void l12_parse_scale_info(const uint8_t *hdr, unsigned kbps_total, Reader &br, Info *sci, ScfOut scf_out)
{
    L12Alloc a;
    l12_alloc_table(hdr, kbps_total, &a);
    sci->total_bands = (uint8_t)a.total_bands;
    sci->stereo_bands = (uint8_t)a.stereo_bands;
    uint8_t scfcod[64];
    int range = 0, left = 0, width = 0, row = 0;
    for (int i = 0; i < a.total_bands; i++)
    {
        if (!left) { left = a.r[range].count; width = a.r[range].width; row = a.r[range].row_off; range++; }
        left--;
        int ba = l12_ba(row, (int)br.get(width), a.layer1);
        sci->bitalloc[2*i] = (uint8_t)ba;
        if (i < a.stereo_bands) ba = l12_ba(row, (int)br.get(width), a.layer1);
        sci->bitalloc[2*i + 1] = (uint8_t)(a.stereo_bands ? ba : 0);
    }
    for (int i = 0; i < 2*a.total_bands; i++)
        scfcod[i] = (uint8_t)(sci->bitalloc[i] ? (a.layer1 ? 2 : (int)br.get(2)) : 6);
    for (int i = 0; i < 2*a.total_bands; i++)
    {
        const int ba = sci->bitalloc[i];
        const int mask = ba ? 4 + ((19 >> scfcod[i]) & 3) : 0;
        int cur = -1;       /* a part without an index of its own keeps the previous part's */
        for (int m = 4, k = 0; m; m >>= 1, k++)
        {
            if (mask & m) cur = (int)br.get(6);
            scf_out(3*i + k, cur, ba);
        }
    }
    for (int i = a.stereo_bands; i < a.total_bands; i++) sci->bitalloc[2*i + 1] = 0;
}

template <class Reader>
MP3_HD void l12_read_scale_info(const uint8_t *hdr, unsigned kbps_total, Reader &br, L12Info *sci, bool want_scf)
{
    l12_parse_scale_info(hdr, kbps_total, br, sci, [&](int idx, int code, int ba) {
        if (want_scf) sci->scf[idx] = code < 0 ? 0.f : l12_scf_value(code, ba);
    });
}

/* the same parse in compact form: what the scale-info kernel leaves in HBM for the sample kernel (272 bytes per frame) */
struct alignas(16) L12Compact
{
    uint8_t bitalloc[64];
    uint8_t scf_index[192];      /* per (band-channel, part); 0xFF: none */
    uint64_t sample_bit_off;     /* absolute bit offset of the first sub-block of samples */
    int32_t sub_bits;            /* bits per sub-block */
    uint8_t total_bands, stereo_bands, group, pad;
};
static_assert(sizeof(L12Compact) == 272, "17 x 16 bytes");

/* allocation that scales (band-channel i): the second channel of a joint-stereo band is scaled with the first one's
 * allocation (its own entry is cleared once the scalefactors are read, minimp3.h:375, 402-405) */
MP3_HD int l12_scale_alloc(const uint8_t *bitalloc, int stereo_bands, int i)
{
    return ((i & 1) && (i >> 1) >= stereo_bands) ? bitalloc[i - 1] : bitalloc[i];
}

/* bits one L12_dequantize_granule sub-block (one value of j) consumes */
template <class Info>
MP3_HD int l12_subblock_bits(const Info &sci, int group_size)
{
    int bits = 0;
    for (int i = 0; i < 2*sci.total_bands; i++)
    {
        const int ba = sci.bitalloc[i];
        if (!ba) continue;
        if (ba < 17) bits += ba*group_size;
        else
        {
            const int mod = (2 << (ba - 17)) + 1;
            bits += mod + 2 - (mod >> 3);
        }
    }
    return bits;
}

/*
 * One band of one sub-block (minimp3.h:434-467, the body of the inner j loop) + scaling and joint-stereo copy
 * (minimp3.h:469-481): group_size consecutive time slots.  v[ch][k] receives sample * scalefactor (zeros beyond
 * total_bands); igr selects which of the three scalefactors of the band applies.
 */
template <class Reader>
MP3_HD void l12_dequant_band(Reader &br, const L12Info &sci, int group_size, int igr, int b, float v[2][3])
{
    for (int ch = 0; ch < 2; ch++)
        for (int k = 0; k < 3; k++) v[ch][k] = 0.f;
    if (b >= sci.total_bands) return;
    for (int ch = 0; ch < 2; ch++)
    {
        const int ba = sci.bitalloc[2*b + ch];
        if (!ba) continue;
        if (ba < 17)
        {
            const int half = (1 << (ba - 1)) - 1;
            for (int k = 0; k < group_size; k++) v[ch][k] = (float)((int)br.get(ba) - half);
        } else
        {
            const unsigned mod = (2u << (ba - 17)) + 1;
            unsigned code = br.get((int)(mod + 2 - (mod >> 3)));
            for (int k = 0; k < group_size; k++, code /= mod) v[ch][k] = (float)((int)(code % mod) - (int)(mod/2));
        }
    }
    if (b >= sci.stereo_bands)
        for (int k = 0; k < 3; k++) v[1][k] = v[0][k];     /* joint stereo: one set of samples, two scalefactors */
    const float s0 = sci.scf[3*(2*b) + igr], s1 = sci.scf[3*(2*b + 1) + igr];
    for (int k = 0; k < 3; k++) { v[0][k] = MP3_FMUL(v[0][k], s0); v[1][k] = MP3_FMUL(v[1][k], s1); }
}

}  // namespace mp3b
#endif
