/*
 * oracle/mp3_oracle.c -- TEST INFRASTRUCTURE ONLY: a plain-C, single-threaded CPU restatement of the
 * decode path of lieff/minimp3 (Layer III, and Layer I/II from the ISO allocation tables), used as the checker for the CUDA kernels.  Never linked into,
 * imported by or executed from the product (minimp3_b200/).
 *
 * Pin status: PINNED.  tests/test_oracle.py checks this file against (a) the golden manifest produced by the
 * UNMODIFIED reference compiled from /root/reference (tests/golden/manifest.json: frame and sample counts for
 * every bundled Layer III vector), (b) the ISO conformance PCM shipped with the reference (PSNR >= 96 dB,
 * the criterion of minimp3_test.c:417-427) and (c), when oracle/_ref is present, the reference's own stage
 * taps: integer spectra, scalefactors and requantised lines bit-exact, PCM within +-1 LSB.
 *
 * How it restates the algorithm (every function cites the reference lines it follows, paths relative to
 * /root/reference): integer stages (sync, side info, reservoir, scalefactors, Huffman) follow the reference
 * exactly, with an INDEPENDENT Huffman decoder (bit-serial match against the ISO code list instead of the
 * packed LUT); the requantiser repeats the reference's float operation sequence (it is part of the format as
 * minimp3 defines it); the filterbank stages (IMDCT, DCT-II matrixing, 512-tap window) are evaluated from the
 * ISO closed forms in DOUBLE precision with dense O(N^2) sums -- deliberately not the fast factorizations --
 * but with the reference's state formulation (un-windowed folded IMDCT tail).
 */
#include "mp3_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "mp3_tables.h"

#define ORC_PI 3.14159265358979323846

/* ------------------------------------------------------------------ bits (minimp3.h:241-262) */
typedef struct { const uint8_t *buf; int pos, limit; } orc_bs;

static uint32_t orc_get_bits(orc_bs *bs, int n)
{
    uint32_t v = 0;
    int i, p = bs->pos;
    bs->pos += n;
    if (bs->pos > bs->limit) return 0;
    for (i = 0; i < n; i++, p++) v = (v << 1) | ((bs->buf[p >> 3] >> (7 - (p & 7))) & 1u);
    return v;
}

/* raw bit at absolute position, no limit (L3_huffman reads through a raw pointer, minimp3.h:765-776) */
static int orc_bit_at(const uint8_t *buf, long p) { return (buf[p >> 3] >> (7 - (p & 7))) & 1; }

/* ------------------------------------------------------------------ header (minimp3.h:61-78, 264-314) */
static int orc_hdr_valid(const uint8_t *h)
{
    return h[0] == 0xff && ((h[1] & 0xF0) == 0xf0 || (h[1] & 0xFE) == 0xe2) && ((h[1] >> 1) & 3) != 0 &&
           (h[2] >> 4) != 15 && ((h[2] >> 2) & 3) != 3;
}
static int orc_hdr_compare(const uint8_t *a, const uint8_t *b)
{
    return orc_hdr_valid(b) && ((a[1] ^ b[1]) & 0xFE) == 0 && ((a[2] ^ b[2]) & 0x0C) == 0 &&
           !(((a[2] & 0xF0) == 0) ^ ((b[2] & 0xF0) == 0));
}
static int orc_mpeg1(const uint8_t *h) { return (h[1] & 8) != 0; }
static int orc_mono(const uint8_t *h) { return (h[3] & 0xC0) == 0xC0; }
static int orc_layer(const uint8_t *h) { return 4 - ((h[1] >> 1) & 3); }
static unsigned orc_kbps(const uint8_t *h)
{
    static const uint16_t m1[3][15] = { { 0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320 },
                                        { 0, 32, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 384 },
                                        { 0, 32, 64, 96, 128, 160, 192, 224, 256, 288, 320, 352, 384, 416, 448 } };
    static const uint16_t m2[3][15] = { { 0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160 },
                                        { 0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160 },
                                        { 0, 32, 48, 56, 64, 80, 96, 112, 128, 144, 160, 176, 192, 224, 256 } };
    int row = ((h[1] >> 1) & 3) - 1;
    return orc_mpeg1(h) ? m1[row][h[2] >> 4] : m2[row][h[2] >> 4];
}
static unsigned orc_hz(const uint8_t *h)
{
    static const unsigned base[3] = { 44100, 48000, 32000 };
    return base[(h[2] >> 2) & 3] >> (!orc_mpeg1(h)) >> (!(h[1] & 0x10));
}
static unsigned orc_frame_samples(const uint8_t *h) { return (h[1] & 6) == 6 ? 384 : (1152 >> ((h[1] & 14) == 2)); }
static int orc_frame_bytes(const uint8_t *h, int free_format)
{
    int n = (int)(orc_frame_samples(h)*orc_kbps(h)*125/orc_hz(h));
    if ((h[1] & 6) == 6) n &= ~3;
    return n ? n : free_format;
}
static int orc_padding(const uint8_t *h) { return (h[2] & 2) ? (((h[1] & 6) == 6) ? 4 : 1) : 0; }

/* minimp3.h:1657-1706 */
static int orc_match(const uint8_t *hdr, int bytes, int frame_bytes)
{
    int i = 0, n;
    for (n = 0; n < 10; n++)
    {
        i += orc_frame_bytes(hdr + i, frame_bytes) + orc_padding(hdr + i);
        if (i + 4 > bytes) return n > 0;
        if (!orc_hdr_compare(hdr, hdr + i)) return 0;
    }
    return 1;
}
static int orc_find_frame(const uint8_t *mp3, int bytes, int *free_format, int *frame_size)
{
    int i, k;
    for (i = 0; i < bytes - 4; i++, mp3++)
    {
        if (!orc_hdr_valid(mp3)) continue;
        {
            int fb = orc_frame_bytes(mp3, *free_format), fp = fb + orc_padding(mp3);
            for (k = 4; !fb && k < 2304 && i + 2*k < bytes - 4; k++)
                if (orc_hdr_compare(mp3, mp3 + k))
                {
                    int f = k - orc_padding(mp3), nf = f + orc_padding(mp3 + k);
                    if (i + k + nf + 4 > bytes || !orc_hdr_compare(mp3, mp3 + k + nf)) continue;
                    fp = k; fb = f; *free_format = f;
                }
            if ((fb && i + fp <= bytes && orc_match(mp3, bytes - i, fb)) || (!i && fp == bytes))
            {
                *frame_size = fp;
                return i;
            }
            *free_format = 0;
        }
    }
    *frame_size = 0;
    return bytes;
}

/* ------------------------------------------------------------------ side info (minimp3.h:484-607) */
typedef struct
{
    int part_23_length, big_values, scalefac_compress, global_gain, block_type, mixed, n_long_sfb, n_short_sfb;
    int table_select[3], region_count[3], subblock_gain[3], preflag, scalefac_scale, count1_table, scfsi;
    const uint8_t *sfb;
} orc_gr;

static int orc_sr_class(const uint8_t *h)
{
    int v = ((h[2] >> 2) & 3) + (((h[1] >> 3) & 1) + ((h[1] >> 4) & 1))*3;
    return v - (v != 0);
}

static int orc_side_info(orc_bs *bs, orc_gr *gr, const uint8_t *hdr)
{
    unsigned tables, scfsi = 0;
    int mdb, sum = 0, n = orc_mono(hdr) ? 1 : 2, sr = orc_sr_class(hdr), e;
    if (orc_mpeg1(hdr)) { n *= 2; mdb = (int)orc_get_bits(bs, 9); scfsi = orc_get_bits(bs, 7 + n); }
    else mdb = (int)(orc_get_bits(bs, 8 + n) >> n);
    for (e = 0; e < n; e++, gr++)
    {
        if (orc_mono(hdr)) scfsi <<= 4;
        gr->part_23_length = (int)orc_get_bits(bs, 12); sum += gr->part_23_length;
        gr->big_values = (int)orc_get_bits(bs, 9);
        if (gr->big_values > 288) return -1;
        gr->global_gain = (int)orc_get_bits(bs, 8);
        gr->scalefac_compress = (int)orc_get_bits(bs, orc_mpeg1(hdr) ? 4 : 9);
        gr->sfb = MP3T_SFB_LONG + sr*23; gr->n_long_sfb = 22; gr->n_short_sfb = 0;
        gr->subblock_gain[0] = gr->subblock_gain[1] = gr->subblock_gain[2] = 0;
        if (orc_get_bits(bs, 1))
        {
            gr->block_type = (int)orc_get_bits(bs, 2);
            if (!gr->block_type) return -1;
            gr->mixed = (int)orc_get_bits(bs, 1);
            gr->region_count[0] = 7; gr->region_count[1] = 255; gr->region_count[2] = 0;
            if (gr->block_type == 2)
            {
                scfsi &= 0x0F0F;
                if (!gr->mixed) { gr->region_count[0] = 8; gr->sfb = MP3T_SFB_SHORT + sr*40; gr->n_long_sfb = 0; gr->n_short_sfb = 39; }
                else { gr->sfb = MP3T_SFB_MIXED + sr*40; gr->n_long_sfb = orc_mpeg1(hdr) ? 8 : 6; gr->n_short_sfb = 30; }
            }
            tables = orc_get_bits(bs, 10) << 5;
            gr->subblock_gain[0] = (int)orc_get_bits(bs, 3);
            gr->subblock_gain[1] = (int)orc_get_bits(bs, 3);
            gr->subblock_gain[2] = (int)orc_get_bits(bs, 3);
        } else
        {
            gr->block_type = 0; gr->mixed = 0;
            tables = orc_get_bits(bs, 15);
            gr->region_count[0] = (int)orc_get_bits(bs, 4);
            gr->region_count[1] = (int)orc_get_bits(bs, 3);
            gr->region_count[2] = 255;
        }
        gr->table_select[0] = (int)(tables >> 10); gr->table_select[1] = (int)((tables >> 5) & 31); gr->table_select[2] = (int)(tables & 31);
        gr->preflag = orc_mpeg1(hdr) ? (int)orc_get_bits(bs, 1) : (gr->scalefac_compress >= 500);
        gr->scalefac_scale = (int)orc_get_bits(bs, 1);
        gr->count1_table = (int)orc_get_bits(bs, 1);
        gr->scfsi = (int)((scfsi >> 12) & 15);
        scfsi <<= 4;
    }
    if (sum + bs->pos > bs->limit + mdb*8) return -1;
    return mdb;
}

/* ------------------------------------------------------------------ requantiser pieces (minimp3.h:642-652, 721-740) */
static float orc_ldexp_q2(float y, int e4)
{
    static const float frac[4] = { 9.31322575e-10f, 7.83145814e-10f, 6.58544508e-10f, 5.53767716e-10f };
    int e;
    do
    {
        float step;
        e = e4 < 120 ? e4 : 120;
        step = frac[e & 3]*(float)((1 << 30) >> (e >> 2));
        y = y*step;
    } while ((e4 -= e) > 0);
    return y;
}

static float orc_pow43(int x)
{
    float frac, poly;
    int sign, mult = 256;
    if (x < 129) return MP3T_POW43[16 + x];
    if (x < 1024) { mult = 16; x <<= 3; }
    sign = (2*x) & 64;
    frac = (float)((x & 63) - sign)/(float)((x & ~63) + sign);
    poly = 1.f + frac*((4.f/3) + frac*(2.f/9));
    return MP3T_POW43[16 + ((x + sign) >> 6)]*poly*(float)mult;
}

/* ------------------------------------------------------------------ scalefactors (minimp3.h:609-714) */
static void orc_scalefactors(const uint8_t *hdr, uint8_t *ist_pos, orc_bs *bs, const orc_gr *gr, float *scf, int ch)
{
    /* ISO 11172-3 slen table; 13818-3 2.4.3.2 nr_of_sfb_block, rows: long, mixed, short */
    static const uint8_t slen[2][16] = { { 0, 0, 0, 0, 3, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4 }, { 0, 1, 2, 3, 0, 1, 2, 3, 1, 2, 3, 1, 2, 3, 2, 3 } };
    static const uint8_t part_m1[3][4] = { { 6, 5, 5, 5 }, { 8, 9, 6, 12 }, { 9, 9, 6, 12 } };
    static const uint8_t part_lsf[6][3][4] = {
        { { 6, 5, 5, 5 }, { 6, 9, 9, 9 }, { 9, 9, 9, 9 } }, { { 6, 5, 7, 3 }, { 6, 9, 12, 6 }, { 9, 9, 12, 6 } },
        { { 11, 10, 0, 0 }, { 15, 18, 0, 0 }, { 18, 18, 0, 0 } }, { { 7, 7, 7, 0 }, { 6, 15, 12, 0 }, { 12, 12, 12, 0 } },
        { { 6, 6, 6, 3 }, { 6, 12, 9, 6 }, { 12, 9, 9, 6 } }, { { 8, 8, 5, 0 }, { 6, 18, 9, 0 }, { 15, 12, 9, 0 } } };
    static const uint8_t pretab[10] = { 1, 1, 1, 1, 2, 2, 3, 3, 3, 2 };
    const int kind = gr->n_short_sfb ? (gr->n_long_sfb ? 1 : 2) : 0;
    const uint8_t *count;
    uint8_t iscf[40 + 3], size[4];
    int i, p, idx = 0, scfsi = gr->scfsi, lsf = !orc_mpeg1(hdr), shift = gr->scalefac_scale + 1, gain_exp;
    float gain;
    if (!lsf)
    {
        size[0] = size[1] = slen[0][gr->scalefac_compress]; size[2] = size[3] = slen[1][gr->scalefac_compress];
        count = part_m1[kind];
    } else
    {
        int ist = (hdr[3] & 0x10) && ch, sfc = gr->scalefac_compress >> ist, tn;
        size[3] = 0;
        if (!ist)
        {
            if (sfc < 400) { tn = 0; size[0] = (sfc >> 4)/5; size[1] = (sfc >> 4)%5; size[2] = (sfc & 15) >> 2; size[3] = sfc & 3; }
            else if (sfc < 500) { tn = 1; sfc -= 400; size[0] = (sfc >> 2)/5; size[1] = (sfc >> 2)%5; size[2] = sfc & 3; }
            else { tn = 2; sfc -= 500; size[0] = sfc/3; size[1] = sfc%3; size[2] = 0; }
        } else
        {
            if (sfc < 180) { tn = 3; size[0] = sfc/36; size[1] = (sfc%36)/6; size[2] = sfc%6; }
            else if (sfc < 244) { tn = 4; sfc -= 180; size[0] = (sfc & 63) >> 4; size[1] = (sfc & 15) >> 2; size[2] = sfc & 3; }
            else { tn = 5; sfc -= 244; size[0] = sfc/3; size[1] = sfc%3; size[2] = 0; }
        }
        count = part_lsf[tn][kind];
        scfsi = 0;
    }
    memset(iscf, 0, sizeof(iscf));
    for (p = 0; p < 4 && count[p]; p++, scfsi *= 2)
    {
        int k, cnt = count[p];
        if (scfsi & 8) memcpy(iscf + idx, ist_pos + idx, cnt);           /* shared with the previous granule */
        else if (!size[p]) { memset(iscf + idx, 0, cnt); memset(ist_pos + idx, 0, cnt); }
        else
        {
            int maxv = lsf ? (1 << size[p]) - 1 : -1;
            for (k = 0; k < cnt; k++)
            {
                int s = (int)orc_get_bits(bs, size[p]);
                ist_pos[idx + k] = (uint8_t)(s == maxv ? -1 : s);
                iscf[idx + k] = (uint8_t)s;
            }
        }
        idx += cnt;
    }
    iscf[idx] = iscf[idx + 1] = iscf[idx + 2] = 0;
    if (gr->n_short_sfb)
    {
        int sh = 3 - shift;
        for (i = 0; i < gr->n_short_sfb; i += 3)
        {
            iscf[gr->n_long_sfb + i + 0] += (uint8_t)(gr->subblock_gain[0] << sh);
            iscf[gr->n_long_sfb + i + 1] += (uint8_t)(gr->subblock_gain[1] << sh);
            iscf[gr->n_long_sfb + i + 2] += (uint8_t)(gr->subblock_gain[2] << sh);
        }
    } else if (gr->preflag)
        for (i = 0; i < 10; i++) iscf[11 + i] += pretab[i];
    gain_exp = gr->global_gain - 4 - 210 - (((hdr[3] & 0xE0) == 0x60) ? 2 : 0);
    gain = orc_ldexp_q2(2048.f, 44 - gain_exp);
    for (i = 0; i < gr->n_long_sfb + gr->n_short_sfb; i++) scf[i] = orc_ldexp_q2(gain, iscf[i] << shift);
}

/* ------------------------------------------------------------------ Huffman (minimp3.h:742-877)
 * Independent decoder: extend the code bit by bit and look it up in the ISO (length, code) list. */
static int orc_huff_pair(const uint8_t *buf, long *pos, int tab, int *x, int *y)
{
    const int n = MP3T_HUFF_CODES_COUNT[tab];
    const mp3t_huffcode_t *c = MP3T_HUFF_CODES + MP3T_HUFF_CODES_BASE[tab];
    uint32_t code = 0;
    int len, k;
    *x = *y = 0;
    if (!n) return 0;
    for (len = 1; len <= 19; len++)
    {
        code = (code << 1) | (uint32_t)orc_bit_at(buf, *pos + len - 1);
        for (k = 0; k < n; k++)
            if (c[k].len == len && c[k].code == code) { *x = c[k].x; *y = c[k].y; *pos += len; return 1; }
    }
    return -1;
}

static void orc_huffman(float *dst, int16_t *q, const uint8_t *buf, long start, const orc_gr *gr, const float *scf, long limit)
{
    static const uint8_t quadA_len[16] = { 1, 4, 4, 5, 4, 6, 5, 6, 4, 5, 5, 6, 5, 6, 6, 6 };
    static const uint8_t quadA_code[16] = { 1, 5, 4, 5, 6, 5, 4, 4, 7, 3, 6, 0, 7, 2, 3, 1 };
    const uint8_t *sfb = gr->sfb;
    long pos = start;
    int big = gr->big_values, ireg = 0, np = 0, line = 0;
    float one = 0.f;
    while (big > 0)
    {
        int tab = gr->table_select[ireg], sfb_cnt = gr->region_count[ireg++], linbits = MP3T_HUFF_LINBITS[tab];
        do
        {
            int pairs;
            np = *sfb++/2;
            pairs = big < np ? big : np;
            one = *scf++;
            do
            {
                int v[2], j;
                orc_huff_pair(buf, &pos, tab, &v[0], &v[1]);
                for (j = 0; j < 2; j++, line++)
                {
                    int m = v[j];
                    if (m == 15 && linbits)
                    {
                        int b;
                        for (b = 0; b < linbits; b++) m += orc_bit_at(buf, pos++) << (linbits - 1 - b);
                    }
                    if (m)
                    {
                        int neg = orc_bit_at(buf, pos++);
                        float mag = one*orc_pow43(m);
                        dst[line] = neg ? -mag : mag;
                        if (q) q[line] = (int16_t)(neg ? -m : m);
                    }
                }
            } while (--pairs);
        } while ((big -= np) > 0 && --sfb_cnt >= 0);
    }
    /* count1 quads (ISO table A: variable length, table B: 4-bit complement) */
    for (np = 1 - big;; line += 4)
    {
        int vwxy = -1, len, s;
        if (gr->count1_table) { int b; vwxy = 0; for (b = 0; b < 4; b++) vwxy = (vwxy << 1) | orc_bit_at(buf, pos + b); vwxy = ~vwxy & 15; len = 4; }
        else
        {
            uint32_t code = 0;
            for (len = 1; len <= 6 && vwxy < 0; len++)
            {
                int k;
                code = (code << 1) | (uint32_t)orc_bit_at(buf, pos + len - 1);
                for (k = 0; k < 16; k++) if (quadA_len[k] == len && quadA_code[k] == code) { vwxy = k; break; }
            }
            len--;
        }
        pos += len;
        if (pos - start > limit) break;
        for (s = 0; s < 4; s++)
        {
            if (!(s & 1)) { if (!--np) { np = *sfb++/2; if (!np) return; one = *scf++; } }
            if (vwxy & (8 >> s))
            {
                int neg = orc_bit_at(buf, pos++);
                dst[line + s] = neg ? -one : one;
                if (q) q[line + s] = (int16_t)(neg ? -1 : 1);
            }
        }
    }
}

/* ------------------------------------------------------------------ stereo (minimp3.h:879-993) */
static void orc_intensity_stereo(float *left, uint8_t *ist_pos, const orc_gr *gr, const uint8_t *hdr)
{
    static const float pan[14] = { 0, 1, 0.21132487f, 0.78867513f, 0.36602540f, 0.63397460f, 0.5f, 0.5f, 0.63397460f, 0.36602540f, 0.78867513f, 0.21132487f, 1, 0 };
    float *right = left + 576;
    int max_band[3] = { -1, -1, -1 }, n_sfb = gr->n_long_sfb + gr->n_short_sfb, blocks = gr->n_short_sfb ? 3 : 1, i, k, start = 0;
    const int mpeg1 = orc_mpeg1(hdr), ms = (hdr[3] & 0x20) != 0, sh = gr[1].scalefac_compress & 1;
    for (i = 0; i < n_sfb; i++)
    {
        for (k = 0; k < gr->sfb[i]; k++) if (right[start + k] != 0) { max_band[i % 3] = i; break; }
        start += gr->sfb[i];
    }
    if (gr->n_long_sfb)
    {
        int m = max_band[0] > max_band[1] ? max_band[0] : max_band[1];
        m = m > max_band[2] ? m : max_band[2];
        max_band[0] = max_band[1] = max_band[2] = m;
    }
    for (i = 0; i < blocks; i++)
    {
        int itop = n_sfb - blocks + i, prev = itop - blocks;
        ist_pos[itop] = (uint8_t)(max_band[i] >= prev ? (mpeg1 ? 3 : 0) : ist_pos[prev]);
    }
    for (i = 0, start = 0; gr->sfb[i]; start += gr->sfb[i], i++)
    {
        unsigned ipos = ist_pos[i];
        if (i > max_band[i % 3] && ipos < (unsigned)(mpeg1 ? 7 : 64))
        {
            float kl, kr, s = ms ? 1.41421356f : 1;
            if (mpeg1) { kl = pan[2*ipos]; kr = pan[2*ipos + 1]; }
            else { kl = 1; kr = orc_ldexp_q2(1, (int)((ipos + 1) >> 1 << sh)); if (ipos & 1) { kl = kr; kr = 1; } }
            kl *= s; kr *= s;
            for (k = 0; k < gr->sfb[i]; k++) { right[start + k] = left[start + k]*kr; left[start + k] = left[start + k]*kl; }
        } else if (ms)
            for (k = 0; k < gr->sfb[i]; k++) { float a = left[start + k], b = right[start + k]; left[start + k] = a + b; right[start + k] = a - b; }
    }
}

/* ------------------------------------------------------------------ hybrid filterbank in double precision */
typedef struct
{
    double tail[2][32][9];       /* un-windowed folded IMDCT tail per subband (minimp3.h:18-23 mdct_overlap) */
    double hist[2][15][64];      /* last 15 matrixed vectors V[64] per channel (ISO 11172-3 annex, fig. A.2) */
} orc_state;

/* reorder (minimp3.h:995-1010) */
static void orc_reorder(float *x, const uint8_t *sfb)
{
    float tmp[576];
    int n = 0, src = 0, i, len;
    for (; (len = *sfb) != 0; sfb += 3, src += 3*len)
        for (i = 0; i < len; i++) { tmp[n++] = x[src + i]; tmp[n++] = x[src + len + i]; tmp[n++] = x[src + 2*len + i]; }
    memcpy(x, tmp, n*sizeof(float));
}

/* antialias (minimp3.h:1012-1045), ISO 11172-3 table B.9 coefficients */
static void orc_antialias(double *x, int nbands)
{
    static const double ci[8] = { -0.6, -0.535, -0.33, -0.185, -0.095, -0.041, -0.0142, -0.0037 };
    int b, i;
    for (b = 0; b < nbands; b++)
        for (i = 0; i < 8; i++)
        {
            double cs = 1.0/sqrt(1.0 + ci[i]*ci[i]), ca = -ci[i]/sqrt(1.0 + ci[i]*ci[i]);   /* reference keeps ca positive */
            double u = x[18*(b + 1) + i], d = x[18*b + 17 - i];
            x[18*(b + 1) + i] = u*cs - d*ca;
            x[18*b + 17 - i] = u*ca + d*cs;
        }
}

/* long block: ISO IMDCT x[n] = sum_k X[k] cos(pi/72 (2n+19)(2k+1)); the reference keeps first half as -x[i]
 * ("sum") and the tail x[18+i], i < 9 (minimp3.h:1087-1142) and windows both with THIS granule's window */
static void orc_imdct_long(const double *X, double *tail, int stop_window, double *out)
{
    double x[36], prev[9];
    int n, k, i;
    for (n = 0; n < 36; n++)
    {
        double s = 0;
        for (k = 0; k < 18; k++) s += X[k]*cos(ORC_PI/72.0*(2*n + 19)*(2*k + 1));
        x[n] = s;
    }
    memcpy(prev, tail, sizeof(prev));
    for (i = 0; i < 9; i++)
    {
        double w0, w1;   /* window applied to (tail, first half) at n = i and mirrored at n = 17 - i */
        if (!stop_window) { w0 = cos(ORC_PI/36.0*(i + 0.5)); w1 = sin(ORC_PI/36.0*(i + 0.5)); }
        else { w0 = i < 6 ? 1.0 : cos(ORC_PI/12.0*(i - 6 + 0.5)); w1 = i < 6 ? 0.0 : sin(ORC_PI/12.0*(i - 6 + 0.5)); }
        out[i] = prev[i]*w0 + x[i]*w1;
        out[17 - i] = prev[i]*w1 + x[17 - i]*w0;
        tail[i] = x[18 + i];
    }
}

/* short block: three 12-point IMDCTs, x_w[n] = sum_k X[3k+w] cos(pi/24 (2n+7)(2k+1)) (minimp3.h:1144-1184) */
static void orc_imdct_short(const double *X, double *tail, double *out)
{
    double xs[3][12], prev[9], u[3];
    int w, n, k, i;
    for (w = 0; w < 3; w++)
        for (n = 0; n < 12; n++)
        {
            double s = 0;
            for (k = 0; k < 6; k++) s += X[3*k + w]*cos(ORC_PI/24.0*(2*n + 7)*(2*k + 1));
            xs[w][n] = s;
        }
    memcpy(prev, tail, sizeof(prev));
    for (i = 0; i < 6; i++) out[i] = prev[i];
    for (i = 0; i < 3; i++) u[i] = prev[6 + i];
    for (w = 0; w < 3; w++)
    {
        double *dst = w == 0 ? out + 6 : (w == 1 ? out + 12 : tail);
        for (i = 0; i < 3; i++)
        {
            double c = cos(ORC_PI/12.0*(i + 0.5)), s = sin(ORC_PI/12.0*(i + 0.5));
            dst[i] = u[i]*c + xs[w][i]*s;
            dst[5 - i] = u[i]*s + xs[w][5 - i]*c;
        }
        for (i = 0; i < 3; i++) u[i] = xs[w][6 + i];
    }
    for (i = 0; i < 3; i++) tail[6 + i] = u[i];
}

/* polyphase synthesis of one time slot (ISO 11172-3 2.4.3.5 + fig. A.2; minimp3.h:1274-1655):
 * V[i] = sum_k cos(pi (16+i)(2k+1)/64) S[k]; out[j] = sum_m D[j+64m] V_{t-2m}[j] + D[j+64m+32] V_{t-2m-1}[32+j],
 * D scaled by 65536 so the result is in int16 units. */
static void orc_synth_slot(double hist[15][64], const double *S, double *out)
{
    double V[64];
    int i, j, k, m;
    for (i = 0; i < 64; i++)
    {
        double s = 0;
        for (k = 0; k < 32; k++) s += cos(ORC_PI*(16 + i)*(2*k + 1)/64.0)*S[k];
        V[i] = s;
    }
    for (j = 0; j < 32; j++)
    {
        double s = 0;
        for (m = 0; m < 8; m++)
        {
            const double *ve = m == 0 ? V : hist[15 - 2*m];       /* slot t - 2m   */
            const double *vo = hist[15 - 2*m - 1];                 /* slot t - 2m-1 */
            s += MP3T_DWIN[64*m + j]*ve[j] + MP3T_DWIN[64*m + 32 + j]*vo[32 + j];
        }
        out[j] = s;
    }
    memmove(hist[0], hist[1], sizeof(double)*14*64);
    memcpy(hist[14], V, sizeof(V));
}

/* minimp3.h:1430-1443 and :1541-1542 */
static int16_t orc_round_scalar(float s)
{
    int16_t v;
    if (s >= 32766.5) return 32767;
    if (s <= -32767.5) return -32768;
    v = (int16_t)(s + .5f);
    v -= (v < 0);
    return v;
}
static int16_t orc_round_simd(float s)
{
    if (s > 32767.0f) s = 32767.0f;
    if (s < -32768.0f) s = -32768.0f;
    return (int16_t)lrintf(s);
}


/* ------------------------------------------------------------------ Layer I / II (minimp3.h:316-482, 1783-1802)
 * Restated from the ISO tables: per subband a bit-allocation field width and the list of quantiser classes it
 * selects (11172-3 tables B.2a-d, 13818-3 table B.1); a class is a number of levels: 2^k - 1 (k-bit codes, one
 * per sample) or 3 / 5 / 9 (three samples packed into one 5 / 7 / 10-bit code). */
typedef struct { int levels, bits, grouped; } orc_qclass;

static orc_qclass orc_l12_class(int levels)
{
    orc_qclass c;
    c.levels = levels; c.grouped = 0; c.bits = 0;
    if (levels == 3) { c.bits = 5; c.grouped = 1; }
    else if (levels == 5) { c.bits = 7; c.grouped = 1; }
    else if (levels == 9) { c.bits = 10; c.grouped = 1; }
    else if (levels) { int k = 0; while ((1 << k) - 1 < levels) k++; c.bits = k; }
    return c;
}

/* levels per allocation code; 0 = nothing transmitted */
static const int ORC_L2_A0[16] = { 0, 3, 7, 15, 31, 63, 127, 255, 511, 1023, 2047, 4095, 8191, 16383, 32767, 65535 };   /* B.2a sb 0-2  */
static const int ORC_L2_A1[16] = { 0, 3, 5, 7, 9, 15, 31, 63, 127, 255, 511, 1023, 2047, 4095, 8191, 65535 };            /* B.2a sb 3-10 */
static const int ORC_L2_A2[8]  = { 0, 3, 5, 7, 9, 15, 31, 65535 };                                                       /* B.2a sb 11-22 */
static const int ORC_L2_A3[4]  = { 0, 3, 5, 65535 };                                                                     /* B.2a sb 23-  */
static const int ORC_L2_C0[16] = { 0, 3, 5, 9, 15, 31, 63, 127, 255, 511, 1023, 2047, 4095, 8191, 16383, 32767 };        /* B.2c sb 0-1, its first 8 / 4 entries for the 3- / 2-bit fields */
static const int ORC_L2_M0[16] = { 0, 3, 5, 7, 9, 15, 31, 63, 127, 255, 511, 1023, 2047, 4095, 8191, 16383 };            /* 13818-3 B.1 sb 0-3 */
static const int ORC_L1[16]    = { 0, 3, 7, 15, 31, 63, 127, 255, 511, 1023, 2047, 4095, 8191, 16383, 32767, 65535 };    /* Layer I: code n -> n+1 bits */

typedef struct
{
    int nbands, stereo_bands;            /* bands with allocation; bands coded separately per channel */
    int nbal[32];                        /* width of the allocation field */
    const int *levels[32];               /* allocation code -> levels */
} orc_l12_layout;

/* which table applies (minimp3.h:317-360) */
static void orc_l12_pick_layout(const uint8_t *hdr, orc_l12_layout *L)
{
    int mode = (hdr[3] >> 6) & 3, sb;
    int stereo_bands = mode == 3 ? 0 : mode == 1 ? (((hdr[3] >> 4) & 3) << 2) + 4 : 32;
    if (orc_layer(hdr) == 1)
    {
        L->nbands = 32;
        for (sb = 0; sb < 32; sb++) { L->nbal[sb] = 4; L->levels[sb] = ORC_L1; }
    } else if (!orc_mpeg1(hdr))
    {
        L->nbands = 30;
        for (sb = 0; sb < 30; sb++)
        {
            L->nbal[sb] = sb < 4 ? 4 : sb < 11 ? 3 : 2;
            L->levels[sb] = sb < 4 ? ORC_L2_M0 : ORC_L2_C0;
        }
    } else
    {
        unsigned kbps = orc_kbps(hdr) >> (mode != 3);
        int sr = (hdr[2] >> 2) & 3;
        if (!kbps) kbps = 192;                       /* free format */
        if (kbps < 56)
        {
            L->nbands = sr == 2 ? 12 : 8;            /* B.2d / B.2c */
            for (sb = 0; sb < L->nbands; sb++) { L->nbal[sb] = sb < 2 ? 4 : 3; L->levels[sb] = ORC_L2_C0; }
        } else
        {
            L->nbands = (kbps >= 96 && sr != 1) ? 30 : 27;   /* B.2b / B.2a */
            for (sb = 0; sb < L->nbands; sb++)
            {
                L->nbal[sb] = sb < 11 ? 4 : sb < 23 ? 3 : 2;
                L->levels[sb] = sb < 3 ? ORC_L2_A0 : sb < 11 ? ORC_L2_A1 : sb < 23 ? ORC_L2_A2 : ORC_L2_A3;
            }
        }
    }
    L->stereo_bands = stereo_bands < L->nbands ? stereo_bands : L->nbands;
}

/* the reference's dequantiser constants (minimp3.h:364-369): 2^-20 * 2^(-r/3) / levels, r = index mod 3, in float */
static float orc_l12_scale(int levels, int index)
{
    static const float third[3] = { 9.53674316e-07f, 7.56931807e-07f, 6.00777173e-07f };
    return third[index % 3]/(float)levels*(float)(1 << 21 >> index/3);
}

/* One Layer I / II frame: bit allocation, scalefactor selection, scalefactors, samples; then each block of 12 time
 * slots goes through the polyphase synthesis.  Returns channel-inclusive samples written, or -1 when the frame
 * reads past its end (the reference drops it and resyncs, minimp3.h:1796-1800). */
static long orc_l12_frame(const uint8_t *hdr, orc_bs *bs, orc_state *st, int nch, long total, int16_t *pcm16, float *pcmf, long pcm_cap)
{
    orc_l12_layout L;
    orc_qclass cls[32][2];
    float sf[32][2][3];
    int scfsi[32][2], sb, ch, part, layer1 = orc_layer(hdr) == 1;
    int slots_per_read = layer1 ? 1 : 3, n_parts = layer1 ? 1 : 3, written = 0;
    static float sample[3][2][32][12];      /* [part][ch][subband][slot] */
    orc_l12_pick_layout(hdr, &L);
    memset(cls, 0, sizeof(cls));
    for (sb = 0; sb < L.nbands; sb++)
    {
        int code = (int)orc_get_bits(bs, L.nbal[sb]), code1 = code;
        if (sb < L.stereo_bands) code1 = (int)orc_get_bits(bs, L.nbal[sb]);
        cls[sb][0] = orc_l12_class(L.levels[sb][code]);
        cls[sb][1] = orc_l12_class(L.stereo_bands ? L.levels[sb][code1] : 0);      /* joint bands: channel 1 shares the allocation */
        if (layer1)
        {   /* Layer I never packs samples: 3 levels are plain 2-bit codes */
            cls[sb][0].grouped = cls[sb][1].grouped = 0;
            if (cls[sb][0].levels == 3) cls[sb][0].bits = 2;
            if (cls[sb][1].levels == 3) cls[sb][1].bits = 2;
        }
    }
    for (sb = 0; sb < L.nbands; sb++)
        for (ch = 0; ch < 2; ch++)
            scfsi[sb][ch] = cls[sb][ch].levels ? (layer1 ? 2 : (int)orc_get_bits(bs, 2)) : -1;
    for (sb = 0; sb < L.nbands; sb++)
        for (ch = 0; ch < 2; ch++)
        {
            /* scfsi: 0 = three scalefactors, 1 = first shared by parts 0,1 + one more, 2 = one for all, 3 = one + one shared by parts 1,2 */
            static const int sent[4][3] = { { 1, 1, 1 }, { 1, 0, 1 }, { 1, 0, 0 }, { 1, 1, 0 } };
            float s = 0;
            for (part = 0; part < 3; part++)
            {
                if (scfsi[sb][ch] >= 0 && sent[scfsi[sb][ch]][part])
                    s = orc_l12_scale(cls[sb][ch].levels, (int)orc_get_bits(bs, 6));
                sf[sb][ch][part] = s;
            }
        }
    /* samples: Layer I 12 reads of one sample, Layer II 3 parts x 4 reads of three samples; subbands in the
     * joint-stereo region are sent once and copied to channel 1 */
    memset(sample, 0, sizeof(sample));
    for (part = 0; part < n_parts; part++)
    {
        int rd, k, n_reads = layer1 ? 12 : 4;
        for (rd = 0; rd < n_reads; rd++)
            for (sb = 0; sb < L.nbands; sb++)
                for (ch = 0; ch < 2; ch++)
                {
                    orc_qclass c = cls[sb][ch];
                    if (!c.levels || (ch == 1 && sb >= L.stereo_bands)) continue;
                    if (c.grouped)
                    {
                        unsigned code = orc_get_bits(bs, c.bits);
                        for (k = 0; k < slots_per_read; k++, code /= (unsigned)c.levels)
                            sample[part][ch][sb][rd*slots_per_read + k] = (float)((int)(code % (unsigned)c.levels) - c.levels/2);
                    } else
                        for (k = 0; k < slots_per_read; k++)
                            sample[part][ch][sb][rd*slots_per_read + k] = (float)((int)orc_get_bits(bs, c.bits) - ((1 << (c.bits - 1)) - 1));
                }
        if (bs->pos > bs->limit) return -1;
    }
    for (part = 0; part < n_parts; part++)
    {
        int slot;
        for (ch = 0; ch < nch; ch++)
            for (slot = 0; slot < 12; slot++)
            {
                double S[32], out[32];
                int j;
                for (sb = 0; sb < 32; sb++)
                {
                    const int src_ch = (ch == 1 && sb >= L.stereo_bands) ? 0 : ch;
                    /* Layer I keeps its single scalefactor in every part (minimp3.h:371-383) */
                    S[sb] = sb < L.nbands ? (double)(sample[part][src_ch][sb][slot]*sf[sb][ch][layer1 ? 2 : part]) : 0.0;
                }
                orc_synth_slot(st->hist[ch], S, out);
                for (j = 0; j < 32; j++)
                {
                    long idx = total + written + (long)(slot*32 + j)*nch + ch;
                    if (idx < pcm_cap)
                    {
                        if (pcm16) pcm16[idx] = (j == 0 || j == 16) ? orc_round_scalar((float)out[j]) : orc_round_simd((float)out[j]);
                        if (pcmf) pcmf[idx] = (float)out[j]*(1.0f/32768.0f);
                    }
                }
            }
        written += 384*nch;
    }
    return written;
}

/* ------------------------------------------------------------------ frame loop (minimp3.h:1713-1806) */
long orc_decode_stream(const uint8_t *buf, long size, int16_t *pcm16, float *pcmf, long pcm_cap, orc_taps *taps, long *n_gc_out)
{
    uint8_t header[4] = { 0, 0, 0, 0 };
    int free_format = 0, reserv = 0;
    uint8_t *maindata = (uint8_t *)calloc(1, 511 + 2304 + 64);
    uint8_t reserv_buf[511];
    orc_state *st = (orc_state *)calloc(1, sizeof(orc_state));
    uint8_t ist_pos[2][40];
    long pos = 0, total = 0, ngc = 0;
    int frame_bytes_out;
    memset(ist_pos, 0, sizeof(ist_pos));
    do
    {
        const uint8_t *mp3 = buf + pos, *hdr;
        int bytes = (int)(size - pos < 0x7fffffff ? size - pos : 0x7fffffff), i = 0, frame_size = 0, nch, mdb, have, success, igr, payload;
        orc_bs bs;
        orc_gr gr[4];
        frame_bytes_out = 0;
        if (bytes > 4 && header[0] == 0xff && orc_hdr_compare(header, mp3))
        {
            frame_size = orc_frame_bytes(mp3, free_format) + orc_padding(mp3);
            if (frame_size != bytes && (frame_size + 4 > bytes || !orc_hdr_compare(mp3, mp3 + frame_size))) frame_size = 0;
        }
        if (!frame_size)
        {
            memset(st, 0, sizeof(*st)); memset(header, 0, 4); reserv = 0; free_format = 0;
            i = orc_find_frame(mp3, bytes, &free_format, &frame_size);
            if (!frame_size || i + frame_size > bytes) { frame_bytes_out = i; pos += i; continue; }
        }
        hdr = mp3 + i;
        memcpy(header, hdr, 4);
        frame_bytes_out = i + frame_size;
        pos += frame_bytes_out;
        nch = orc_mono(hdr) ? 1 : 2;
        bs.buf = hdr + 4; bs.pos = 0; bs.limit = (frame_size - 4)*8;
        if (!(hdr[1] & 1)) orc_get_bits(&bs, 16);
        if (orc_layer(hdr) != 3)
        {
            long n = orc_l12_frame(hdr, &bs, st, nch, total, pcm16, pcmf, pcm_cap);
            if (n < 0) header[0] = 0;      /* overrun: frame dropped, decoder resyncs from scratch */
            else total += n;
            continue;
        }
        mdb = orc_side_info(&bs, gr, hdr);
        if (mdb < 0 || bs.pos > bs.limit) { header[0] = 0; continue; }
        /* reservoir (minimp3.h:1212-1236) */
        payload = (bs.limit - bs.pos)/8;
        have = reserv < mdb ? reserv : mdb;
        memcpy(maindata, reserv_buf + (reserv - mdb > 0 ? reserv - mdb : 0), have);
        memcpy(maindata + have, bs.buf + bs.pos/8, payload);
        success = reserv >= mdb;
        {
            long bitpos = 0;
            if (success)
                for (igr = 0; igr < (orc_mpeg1(hdr) ? 2 : 1); igr++)
                {
                    float grbuf[2][576], scf[40];
                    double xd[576], sub[2][576];
                    int ch, sb, t, k;
                    memset(grbuf, 0, sizeof(grbuf));
                    for (ch = 0; ch < nch; ch++)
                    {
                        orc_gr *g = gr + igr*nch + ch;
                        orc_bs mb;
                        int16_t *q = taps && taps->q && ngc + ch < taps->cap ? taps->q + (ngc + ch)*576 : 0;
                        mb.buf = maindata; mb.pos = (int)bitpos; mb.limit = (have + payload)*8;
                        memset(scf, 0, sizeof(scf));
                        orc_scalefactors(hdr, ist_pos[ch], &mb, g, scf, ch);
                        if (q) memset(q, 0, 576*sizeof(int16_t));
                        orc_huffman(grbuf[ch], q, maindata, mb.pos, g, scf, bitpos + g->part_23_length - mb.pos);
                        if (taps && ngc + ch < taps->cap)
                        {
                            if (taps->scf) memcpy(taps->scf + (ngc + ch)*40, scf, 40*sizeof(float));
                            if (taps->xr) memcpy(taps->xr + (ngc + ch)*576, grbuf[ch], 576*sizeof(float));
                        }
                        bitpos += g->part_23_length;
                    }
                    if ((hdr[3] & 0x10) && nch == 2) orc_intensity_stereo(grbuf[0], ist_pos[1], gr + igr*nch, hdr);
                    else if ((hdr[3] & 0xE0) == 0x60)
                        for (k = 0; k < 576; k++) { float a = grbuf[0][k], b = grbuf[1][k]; grbuf[0][k] = a + b; grbuf[1][k] = a - b; }
                    for (ch = 0; ch < nch; ch++)
                    {
                        orc_gr *g = gr + igr*nch + ch;
                        int n_long = (g->mixed ? 2 : 0) << (orc_sr_class(hdr) == 1);   /* 8 kHz MPEG-2.5: 4 subbands (minimp3.h:1260) */
                        int aa = 31;
                        if (g->n_short_sfb) { aa = n_long - 1; orc_reorder(grbuf[ch] + n_long*18, g->sfb + g->n_long_sfb); }
                        for (k = 0; k < 576; k++) xd[k] = grbuf[ch][k];
                        orc_antialias(xd, aa);
                        for (sb = 0; sb < 32; sb++)
                        {
                            double o[18];
                            if (g->block_type == 2 && sb >= n_long) orc_imdct_short(xd + 18*sb, st->tail[ch][sb], o);
                            else orc_imdct_long(xd + 18*sb, st->tail[ch][sb], g->block_type == 3 && sb >= n_long, o);
                            for (t = 0; t < 18; t++) sub[ch][18*sb + t] = ((sb & 1) && (t & 1)) ? -o[t] : o[t];   /* minimp3.h:1186-1192 */
                        }
                        if (taps && taps->sub && ngc + ch < taps->cap)
                            for (k = 0; k < 576; k++) taps->sub[(ngc + ch)*576 + k] = (float)sub[ch][k];
                    }
                    for (ch = 0; ch < nch; ch++)
                        for (t = 0; t < 18; t++)
                        {
                            double S[32], out[32];
                            int j;
                            for (sb = 0; sb < 32; sb++) S[sb] = sub[ch][18*sb + t];
                            orc_synth_slot(st->hist[ch], S, out);
                            for (j = 0; j < 32; j++)
                            {
                                long idx = total + (long)(t*32 + j)*nch + ch;
                                if (idx < pcm_cap)
                                {
                                    if (pcm16) pcm16[idx] = (j == 0 || j == 16) ? orc_round_scalar((float)out[j]) : orc_round_simd((float)out[j]);
                                    if (pcmf) pcmf[idx] = (float)out[j]*(1.0f/32768.0f);
                                }
                            }
                        }
                    total += 576*nch;
                    ngc += nch;
                }
            {
                int p = (int)((success ? bitpos : 0) + 7)/8, remains = have + payload - p;
                if (remains > 511) { p += remains - 511; remains = 511; }
                if (remains > 0) memmove(reserv_buf, maindata + p, remains);
                reserv = remains;
            }
        }
    } while (frame_bytes_out);
    free(maindata); free(st);
    if (n_gc_out) *n_gc_out = ngc;
    return total;
}
