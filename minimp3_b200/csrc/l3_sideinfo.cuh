/*
 * l3_sideinfo.cuh -- Layer III side info -> granule-channel descriptors, host and device.
 *
 * Restates L3_read_side_info (/root/reference/minimp3.h:484-607).  The host walker (host_parse.cpp) uses the
 * "lite" reader: only what frame sync, the bit reservoir and the lane order need (main_data_begin, the lengths,
 * the validity checks).  The full parse and the GcDesc records are produced on the GPU by k_l3_side_info
 * (k_sideinfo.cu) from the raw side-info bytes -- and by the same code on the host for the single-frame API
 * and the CPU test harness.
 */
#ifndef MP3B_L3_SIDEINFO_CUH
#define MP3B_L3_SIDEINFO_CUH
#include <stdint.h>
#include <string.h>

#include "l3_types.h"

#if defined(__CUDACC__)
#define MP3S_HD __host__ __device__ __forceinline__
#else
#define MP3S_HD inline
#endif

namespace mp3b {

/* One parsed granule-channel of side info. */
struct GranuleSide
{
    uint16_t part_23_length, big_values, scalefac_compress;
    uint8_t global_gain, block_type, mixed_block_flag, n_long_sfb, n_short_sfb;
    uint8_t table_select[3], region_count[3], subblock_gain[3];
    uint8_t preflag, scalefac_scale, count1_table, scfsi;
};

/* MSB-first reader over a small byte window (the side info is at most 32 bytes); fields of at most 25 bits */
struct SiBits
{
    const uint8_t *p;
    int pos, limit;
    MP3S_HD void init(const uint8_t *data, int bytes) { p = data; pos = 0; limit = bytes*8; }
    MP3S_HD uint32_t get(int n)
    {
        const int s = pos;
        pos += n;
        if (pos > limit || n == 0) return 0;
        uint32_t acc = 0;
        int b = s >> 3;
        const int last = (pos - 1) >> 3;
        uint64_t w = 0;
        for (; b <= last; b++) w = (w << 8) | p[b];
        acc = (uint32_t)((w >> (7 - ((pos - 1) & 7))) & ((1ull << n) - 1));
        return acc;
    }
};

MP3S_HD bool si_hdr_mpeg1(const uint8_t *h) { return (h[1] & 0x8) != 0; }
MP3S_HD bool si_hdr_mono(const uint8_t *h) { return (h[3] & 0xC0) == 0xC0; }

/* row of the sfb tables: MPEG-2.5: 0 (11025, 12000), 1 (8000); MPEG-2: 2, 3, 4; MPEG-1: 5, 6, 7 */
MP3S_HD int si_sample_rate_class(const uint8_t *hdr)
{
    const int version_steps = ((hdr[1] >> 3) & 1) + ((hdr[1] >> 4) & 1);
    const int idx = ((hdr[2] >> 2) & 3) + version_steps*3;
    return idx - (idx != 0);
}

/* Full parse (minimp3.h:484-607).  BR: any reader with get(n), pos, limit.  Returns main_data_begin or -1.
 * BUDGET = false skips the last check (main data longer than the frame plus its reservoir reach), which needs the
 * frame size: the device reader only sees the side-info bytes, and the walker has applied the check already. */
template <bool BUDGET, class BR>
MP3S_HD int read_side_info_t(BR *bs, GranuleSide *gr, const uint8_t *hdr)
{
    const bool mpeg1 = si_hdr_mpeg1(hdr), mono = si_hdr_mono(hdr);
    int entries = mono ? 1 : 2;
    unsigned scfsi = 0;
    int main_data_begin, part_23_sum = 0;

    if (mpeg1)
    {
        entries *= 2;
        main_data_begin = (int)bs->get(9);
        /* private bits are read together with scfsi and leak into granule 0 (quirk kept on purpose) */
        scfsi = bs->get(7 + entries);
    } else
    {
        main_data_begin = (int)(bs->get(8 + entries) >> entries);
    }

    for (int e = 0; e < entries; e++, gr++)
    {
        if (mono) scfsi <<= 4;
        gr->part_23_length = (uint16_t)bs->get(12);
        part_23_sum += gr->part_23_length;
        gr->big_values = (uint16_t)bs->get(9);
        if (gr->big_values > 288) return -1;
        gr->global_gain = (uint8_t)bs->get(8);
        gr->scalefac_compress = (uint16_t)bs->get(mpeg1 ? 4 : 9);
        gr->n_long_sfb = 22;
        gr->n_short_sfb = 0;
        unsigned tables;
        if (bs->get(1))   /* window switching */
        {
            gr->block_type = (uint8_t)bs->get(2);
            if (!gr->block_type) return -1;
            gr->mixed_block_flag = (uint8_t)bs->get(1);
            gr->region_count[0] = 7;
            gr->region_count[1] = 255;
            gr->region_count[2] = 0;
            if (gr->block_type == 2)
            {
                scfsi &= 0x0F0F;
                if (!gr->mixed_block_flag)
                {
                    gr->region_count[0] = 8;
                    gr->n_long_sfb = 0;
                    gr->n_short_sfb = 39;
                } else
                {
                    gr->n_long_sfb = mpeg1 ? 8 : 6;
                    gr->n_short_sfb = 30;
                }
            }
            tables = bs->get(10) << 5;
            gr->subblock_gain[0] = (uint8_t)bs->get(3);
            gr->subblock_gain[1] = (uint8_t)bs->get(3);
            gr->subblock_gain[2] = (uint8_t)bs->get(3);
        } else
        {
            gr->block_type = 0;
            gr->mixed_block_flag = 0;
            tables = bs->get(15);
            gr->region_count[0] = (uint8_t)bs->get(4);
            gr->region_count[1] = (uint8_t)bs->get(3);
            gr->region_count[2] = 255;
            gr->subblock_gain[0] = gr->subblock_gain[1] = gr->subblock_gain[2] = 0;
        }
        gr->table_select[0] = (uint8_t)(tables >> 10);
        gr->table_select[1] = (uint8_t)((tables >> 5) & 31);
        gr->table_select[2] = (uint8_t)(tables & 31);
        gr->preflag = mpeg1 ? (uint8_t)bs->get(1) : (uint8_t)(gr->scalefac_compress >= 500);
        gr->scalefac_scale = (uint8_t)bs->get(1);
        gr->count1_table = (uint8_t)bs->get(1);
        gr->scfsi = (uint8_t)((scfsi >> 12) & 15);
        scfsi <<= 4;
    }
    if (BUDGET && part_23_sum + bs->pos > bs->limit + main_data_begin*8) return -1;
    return main_data_begin;
}

/* What the walker needs, at a third of the reads: main_data_begin, part_23_length (reservoir accounting, lane
 * order) and the three conditions under which the reference rejects the frame.  Leaves bs->pos where the full
 * parse leaves it.  gr[].part_23_length / big_values / block_type are the only fields written. */
template <class BR>
MP3S_HD int read_side_info_lite_t(BR *bs, GranuleSide *gr, const uint8_t *hdr)
{
    const bool mpeg1 = si_hdr_mpeg1(hdr), mono = si_hdr_mono(hdr);
    int entries = mono ? 1 : 2;
    int main_data_begin, part_23_sum = 0;
    if (mpeg1)
    {
        entries *= 2;
        main_data_begin = (int)bs->get(9);
        bs->pos += 7 + entries;
    } else
        main_data_begin = (int)(bs->get(8 + entries) >> entries);
    for (int e = 0; e < entries; e++, gr++)
    {
        const uint32_t v = bs->get(21);              /* part_23_length:12 big_values:9 */
        gr->part_23_length = (uint16_t)(v >> 9);
        gr->big_values = (uint16_t)(v & 511);
        part_23_sum += gr->part_23_length;
        if (gr->big_values > 288) return -1;
        bs->pos += 8 + (mpeg1 ? 4 : 9);              /* global_gain, scalefac_compress */
        const uint32_t w = bs->get(3);               /* window_switching_flag:1 block_type:2 (or table bits) */
        gr->block_type = (w & 4) ? (uint8_t)(w & 3) : (uint8_t)0;
        if ((w & 4) && !(w & 3)) return -1;
        bs->pos += 20 + (mpeg1 ? 3 : 2);             /* rest of the 22 switched / unswitched bits, flags */
    }
    if (part_23_sum + bs->pos > bs->limit + main_data_begin*8) return -1;
    return main_data_begin;
}

/* the header-derived part of a GcDesc, the same for every granule-channel of a frame */
struct HdrBits
{
    uint8_t flags0, flags1, sr_idx;
    MP3S_HD HdrBits(const uint8_t *hdr, int nch)
    {
        const bool ms = (hdr[3] & 0xE0) == 0x60;        /* joint stereo + MS bit (HDR_IS_MS_STEREO) */
        const bool is = (hdr[3] & 0x10) != 0;           /* HDR_TEST_I_STEREO: bit is tested whatever the mode */
        flags0 = (uint8_t)((ms ? GCF_MS_STEREO : 0) | (is ? GCF_I_STEREO : 0));
        flags1 = (uint8_t)((si_hdr_mpeg1(hdr) ? GCF1_MPEG1 : 0) | (nch == 2 ? GCF1_STEREO : 0) |
                           ((ms && !is && nch == 2) ? GCF1_MS_PLAIN : 0) | ((hdr[3] & 0x20) ? GCF1_MS_TEST : 0));
        sr_idx = (uint8_t)si_sample_rate_class(hdr);
    }
};

MP3S_HD void fill_gc_desc_hb(GcDesc *d, const GranuleSide &g, const HdrBits &hb, int ch, int igr, uint64_t bit_off)
{
    d->bit_off = bit_off;
    d->part_23_length = g.part_23_length;
    d->big_values = g.big_values;
    d->scalefac_compress = g.scalefac_compress;
    d->global_gain = g.global_gain;
    d->flags0 = (uint8_t)(hb.flags0 | (g.block_type & 3) | (g.mixed_block_flag ? GCF_MIXED : 0) | (g.preflag ? GCF_PREFLAG : 0) |
                          (g.scalefac_scale ? GCF_SCALEFAC_SCALE : 0) | (g.count1_table ? GCF_COUNT1_TABLE : 0));
    for (int i = 0; i < 3; i++)
    {
        d->table_select[i] = g.table_select[i];
        d->region_count[i] = g.region_count[i];
        d->subblock_gain[i] = g.subblock_gain[i];
    }
    d->scfsi = g.scfsi;
    d->n_long_sfb = g.n_long_sfb;
    d->n_short_sfb = g.n_short_sfb;
    d->sr_idx = hb.sr_idx;
    d->flags1 = (uint8_t)(hb.flags1 | (ch ? GCF1_CH : 0) | (igr ? GCF1_GR1 : 0));
    d->pad[0] = d->pad[1] = 0;
}

/* All descriptors of one frame from its raw side info: gcs[igr*nch + ch], bit offsets running from md_bit.
 * Returns the number of descriptors written (0 if the side info is rejected -- the walker has checked that). */
MP3S_HD int frame_descs(const FrameRec &fr, GcDesc *out)
{
    SiBits bs;
    bs.init(fr.si, 32);
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

}  // namespace mp3b
#endif
