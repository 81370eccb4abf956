/*
 * l3_entropy.cuh -- per-granule-channel entropy decode: part 2 (scalefactors) and part 3 (Huffman big
 * values + count1 quads) with requantisation, written as a __host__ __device__ state machine so that
 * the code one GPU lane executes can also be stepped on the CPU by the unit tests.
 *
 * Restates L3_read_scalefactors / L3_decode_scalefactors (minimp3.h:609-714) and L3_huffman
 * (minimp3.h:742-877).  One call of GranuleDecoder::step() produces the next TWO spectral lines: a
 * big-value pair, or half of a count1 quad -- this keeps all lanes of a warp on the same line index.
 */
#ifndef MP3B_L3_ENTROPY_CUH
#define MP3B_L3_ENTROPY_CUH
#include <stdint.h>

#include "l3_math.cuh"
#include "l3_types.h"

namespace mp3b {

/*
 * Table access.  By default through the pointers below.  The entropy kernel keeps the tables in a file-scope
 * __shared__ struct and defines these accessors (before including this header) to name that struct directly:
 * a pointer that has been through a struct member is a generic address to the compiler, and every use of one
 * costs it the shared-window base again (S2R + LEA) -- a dozen extra instructions per decoded line pair.
 */
#ifndef MP3B_ENT_LUT
#define MP3B_ENT_LUT(T, i) ((T).lut[i])
#define MP3B_ENT_POW43(T) ((T).pow43)
#define MP3B_ENT_TABINFO(T, i) ((T).tabinfo[i])
#define MP3B_ENT_SFB(T, kind, row) \
    ((kind) == 0 ? (T).sfb_long + (row)*23 : (kind) == 1 ? (T).sfb_mixed + (row)*40 : (T).sfb_short + (row)*40)
#endif

/* pointers to the decode tables (plain arrays on the host) */
struct EntropyTables
{
    const uint16_t *lut;        /* MP3T_HUFF_LUT */
    const float *pow43;         /* MP3T_POW43 */
    const uint32_t *tabinfo;    /* [32] base | rootw << 16 | linbits << 20 */
    uint32_t c1info[2];         /* count1 table A / B: LUT base | root width << 16 (entries: len << 8 | vwxy) */
    const uint8_t *sfb_long, *sfb_short, *sfb_mixed;
    uint32_t smem_base;     /* device build: shared-space address of the kernel's table block (k_entropy.cu) */
};

/* MSB-first bit reader over big-endian bytes, fed by aligned 32-bit loads.  The 64-bit window is kept as two
 * 32-bit halves (funnel shifts instead of 64-bit shifts) and two more words are fetched ahead of need so the
 * global-load latency overlaps with the decoding of the bits already in the window. */
/* m (>= +0) with the sign bit `sign` (0 or 0x80000000) put on */
MP3_HD float flip_sign(float m, uint32_t sign)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(__float_as_uint(m) | sign);
#else
    union { float f; uint32_t u; } c;
    c.f = m;
    c.u |= sign;
    return c.f;
#endif
}

struct BitRd
{
    const uint32_t *wp;     /* next word to fetch */
    const uint32_t *wend;   /* end of the buffer (host harness only) */
    uint32_t hi, lo;        /* two consecutive words of the stream, big-endian applied */
    uint32_t next, next2;   /* the two words behind them; next2 still in memory byte order */
    int off;                /* cursor: bit offset into hi:lo, < 64; < 32 after normalize() */
    int base;               /* bits consumed up to the first bit of hi */

    /* a word as stored (big-endian bit order not applied yet).  The device buffer carries 64 bytes of slack
     * behind the last payload byte (capi.cu), the host harness is bounds-checked. */
    MP3_HD uint32_t raw(const uint32_t *p) const
    {
#if defined(__CUDA_ARCH__)
        return __ldg(p);
#else
        return p >= wend ? 0u : *p;
#endif
    }
    static MP3_HD uint32_t be(uint32_t v)
    {
#if defined(__CUDA_ARCH__)
        return __byte_perm(v, 0, 0x0123);
#else
        return __builtin_bswap32(v);
#endif
    }
    /* upper 32 bits of (h:l) << n, 0 <= n < 32 */
    static MP3_HD uint32_t shl_hi(uint32_t h, uint32_t l, int n)
    {
#if defined(__CUDA_ARCH__)
        return __funnelshift_l(l, h, n);
#else
        return n ? (h << n) | (l >> (32 - n)) : h;
#endif
    }
    MP3_HD void init(const uint32_t *words, uint64_t nw, uint64_t bit_off)
    {
        wend = words + nw;
        wp = words + (bit_off >> 5);
        hi = be(raw(wp));
        lo = be(raw(wp + 1));
        next = be(raw(wp + 2));
        next2 = raw(wp + 3);
        wp += 4;
        off = (int)(bit_off & 31);
        base = -off;
    }
    /* Brings the cursor back into the first word (needs off < 64: true wherever it is called -- a step without
     * escapes moves the cursor by 21 bits at most, one with escapes calls normalize_wide() itself).  The words only
     * move, nothing is shifted: the cursor is applied by top() / top2() when bits are looked at, and consuming bits
     * is one addition.  Written without a branch: in a warp of 32 granules some lane crosses a word at almost every
     * call. */
    MP3_HD void normalize()
    {
        const int step = off & 32;              /* off < 64: 32 when the cursor has left the first word, else 0 */
        const bool adv = step != 0;
        hi = adv ? lo : hi;
        lo = adv ? next : lo;
        next = adv ? be(next2) : next;          /* the byte swap waits here, one word after the load was issued */
        if (adv) { next2 = raw(wp); wp += 1; }
        off ^= step;
        base += step;
    }
    /* the same for a cursor anywhere below 96 (takes one word off, may leave the cursor in the second word) */
    MP3_HD void normalize_wide()
    {
        const bool adv = off >= 32;
        hi = adv ? lo : hi;
        lo = adv ? next : lo;
        next = adv ? be(next2) : next;
        if (adv) { next2 = raw(wp); wp += 1; off -= 32; base += 32; }
    }
    /* the 32 bits at the cursor, and the 32 behind them (cursor < 32) */
    MP3_HD uint32_t top() const { return shl_hi(hi, lo, off); }
    MP3_HD uint32_t top2() const { return shl_hi(lo, next, off); }
    MP3_HD void skip(int n) { off += n; }
    MP3_HD int bits_consumed() const { return base + off; }
    /* 0 <= n <= 32 */
    MP3_HD uint32_t get(int n)
    {
        normalize();
        const uint32_t v = n ? top() >> (32 - n) : 0u;
        skip(n);
        return v;
    }
};

struct ScfLayout
{
    uint8_t size[4];
    uint8_t count[4];
};

/* partition sizes as four packed bytes (count[0] in the low byte); kind: 0 long, 1 mixed, 2 short.
 * ISO 11172-3 2.4.2.7 (MPEG-1) and 13818-3 2.4.3.2 nr_of_sfb_block (LSF). */
MP3_HD uint32_t scf_partition_mpeg1(int kind)
{
    return kind == 0 ? 0x05050506u : kind == 1 ? 0x0C060908u : 0x0C060909u;
}

MP3_HD uint32_t scf_partition_lsf(int tn, int kind)
{
    switch (tn*3 + kind)
    {
    case 0: return 0x05050506u; case 1: return 0x09090906u; case 2: return 0x09090909u;
    case 3: return 0x03070506u; case 4: return 0x060C0906u; case 5: return 0x060C0909u;
    case 6: return 0x00000A0Bu; case 7: return 0x0000120Fu; case 8: return 0x00001212u;
    case 9: return 0x00070707u; case 10: return 0x000C0F06u; case 11: return 0x000C0C0Cu;
    case 12: return 0x03060606u; case 13: return 0x06090C06u; case 14: return 0x0609090Cu;
    case 15: return 0x00050808u; case 16: return 0x00091206u; default: return 0x00090C0Fu;
    }
}

/* slen / partition selection (minimp3.h:666-687) */
MP3_HD ScfLayout scf_layout(const GcDesc &d)
{
    ScfLayout L;
    const int kind = d.n_short_sfb ? (d.n_long_sfb ? 1 : 2) : 0;
    uint32_t part;
    if (d.flags1 & GCF1_MPEG1)
    {
        /* ISO slen1 / slen2 per scalefac_compress, one nibble each */
        const uint64_t slen1 = 0x4433322211130000ull, slen2 = 0x3232132132103210ull;
        const int sh = (d.scalefac_compress & 15)*4;
        const int s1 = (int)((slen1 >> sh) & 15), s2 = (int)((slen2 >> sh) & 15);
        L.size[0] = L.size[1] = (uint8_t)s1;
        L.size[2] = L.size[3] = (uint8_t)s2;
        part = scf_partition_mpeg1(kind);
    } else
    {
        const bool ist = (d.flags0 & GCF_I_STEREO) && (d.flags1 & GCF1_CH);
        int sfc = d.scalefac_compress >> (ist ? 1 : 0);
        int tn;
        int s0, s1, s2, s3 = 0;
        if (!ist)
        {
            if (sfc < 400)      { tn = 0; s0 = (sfc >> 4)/5; s1 = (sfc >> 4)%5; s2 = (sfc & 15) >> 2; s3 = sfc & 3; }
            else if (sfc < 500) { tn = 1; sfc -= 400; s0 = (sfc >> 2)/5; s1 = (sfc >> 2)%5; s2 = sfc & 3; }
            else                { tn = 2; sfc -= 500; s0 = sfc/3; s1 = sfc%3; s2 = 0; }
        } else
        {
            if (sfc < 180)      { tn = 3; s0 = sfc/36; s1 = (sfc%36)/6; s2 = sfc%6; }
            else if (sfc < 244) { tn = 4; sfc -= 180; s0 = (sfc & 63) >> 4; s1 = (sfc & 15) >> 2; s2 = sfc & 3; }
            else                { tn = 5; sfc -= 244; s0 = sfc/3; s1 = sfc%3; s2 = 0; }
        }
        L.size[0] = (uint8_t)s0; L.size[1] = (uint8_t)s1; L.size[2] = (uint8_t)s2; L.size[3] = (uint8_t)s3;
        part = scf_partition_lsf(tn, kind);
    }
    L.count[0] = (uint8_t)part; L.count[1] = (uint8_t)(part >> 8); L.count[2] = (uint8_t)(part >> 16); L.count[3] = (uint8_t)(part >> 24);
    return L;
}

/* L3_read_scalefactors (minimp3.h:609-640).  dst/prev are byte arrays with element stride `st`.
 * prev == nullptr means "no previous granule": copied partitions read as zero.  Returns bands read. */
MP3_HD int read_scalefactors(BitRd &br, const ScfLayout &L, int scfsi, bool lsf, uint8_t *dst, const uint8_t *prev,
                             int st, uint8_t *ist_out)
{
    int idx = 0;
    /* kept rolled: run once per granule, the unrolled form was 3000 instructions of straight-line code that
     * every warp had to pull through the instruction cache */
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int p = 0; p < 4 && L.count[p]; p++, scfsi *= 2)
    {
        const int cnt = L.count[p];
        const int bits = L.size[p];
        const bool copy = (scfsi & 8) != 0;
        const int max_scf = lsf ? (1 << bits) - 1 : -1;
        if (!copy && bits)
        {
            /* up to six fields (<= 5 bits each) per refill of the window instead of one get() per field */
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            for (int k = 0; k < cnt; )
            {
                br.normalize();
                uint32_t win = br.top();
                const int n = cnt - k < 6 ? cnt - k : 6;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
                for (int i = 0; i < n; i++, k++)
                {
                    const int s = (int)(win >> (32 - bits));
                    win <<= bits;
                    dst[(idx + k)*st] = (uint8_t)s;
                    if (ist_out) ist_out[idx + k] = (uint8_t)((s == max_scf) ? 255 : s);
                }
                br.skip(n*bits);
            }
        } else
        {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            for (int k = 0; k < cnt; k++)
            {
                const int s = (copy && prev) ? prev[(idx + k)*st] : 0;
                dst[(idx + k)*st] = (uint8_t)s;
                if (ist_out) ist_out[idx + k] = (uint8_t)s;
            }
        }
        idx += cnt;
    }
    dst[(idx + 0)*st] = 0;
    dst[(idx + 1)*st] = 0;
    dst[(idx + 2)*st] = 0;
    return idx;
}

enum { EMODE_BIG = 0, EMODE_C1A = 1, EMODE_C1B = 2, EMODE_DONE = 3 };

struct GranuleDecoder
{
    BitRd br;
    const uint8_t *sfbw;      /* band widths of this granule's partition */
    const uint8_t *iscf;      /* final integer scalefactors, stride iscf_st */
    int iscf_st;
    const uint8_t *sfbw_run, *iscf_run;   /* the next band's entries: per-lane pointers that only ever advance */
    float gain, one;
    int scf_shift, limit;
    int mode, big_left, band_left, ireg, reg_left, pending, extent;
    uint32_t tinfo;
    uint8_t table_select[3], region_count[3];
    uint32_t c1info;          /* T.c1info[] entry of this granule's count1 table */

    /* scalefactors + set-up.  iscf_buf / prev_buf: 40 bytes each at stride st.  d0: descriptor of granule 0
     * of the same frame/channel when this is granule 1 with scfsi != 0 (else nullptr).
     * ist_out: optional [40] intensity positions. */
    MP3_HD void init(const GcDesc &d, const GcDesc *d0, const uint32_t *words, uint64_t n_words, const EntropyTables &T,
                     uint8_t *iscf_buf, uint8_t *prev_buf, int st, uint8_t *ist_out, int *n_ist)
    {
        const bool mpeg1 = (d.flags1 & GCF1_MPEG1) != 0;
        const uint8_t *prev = nullptr;
        if (d0)
        {
            /* scfsi: bands shared with granule 0 of the same channel -> re-read that granule's part 2 */
            const ScfLayout L0 = scf_layout(*d0);
            br.init(words, n_words, d0->bit_off);
            read_scalefactors(br, L0, d0->scfsi, false, prev_buf, nullptr, st, nullptr);
            prev = prev_buf;
        }
        const ScfLayout L = scf_layout(d);
        br.init(words, n_words, d.bit_off);
        const int n_read = read_scalefactors(br, L, mpeg1 ? d.scfsi : 0, !mpeg1, iscf_buf, prev, st, ist_out);
        if (n_ist) *n_ist = n_read;
        scf_shift = ((d.flags0 & GCF_SCALEFAC_SCALE) ? 1 : 0) + 1;
        /* subblock gain / preemphasis (minimp3.h:690-706); uint8 arithmetic wraps like the reference's */
        if (d.n_short_sfb)
        {
            const int sh = 3 - scf_shift;
            for (int i = 0; i < d.n_short_sfb; i += 3)
            {
                iscf_buf[(d.n_long_sfb + i + 0)*st] = (uint8_t)(iscf_buf[(d.n_long_sfb + i + 0)*st] + (d.subblock_gain[0] << sh));
                iscf_buf[(d.n_long_sfb + i + 1)*st] = (uint8_t)(iscf_buf[(d.n_long_sfb + i + 1)*st] + (d.subblock_gain[1] << sh));
                iscf_buf[(d.n_long_sfb + i + 2)*st] = (uint8_t)(iscf_buf[(d.n_long_sfb + i + 2)*st] + (d.subblock_gain[2] << sh));
            }
        } else if (d.flags0 & GCF_PREFLAG)
        {
            /* ISO pretab for bands 11..20: 1 1 1 1 2 2 3 3 3 2, two bits each */
            const uint32_t pretab = 0xBFA55u;
            for (int i = 0; i < 10; i++)
                iscf_buf[(11 + i)*st] = (uint8_t)(iscf_buf[(11 + i)*st] + ((pretab >> (2*i)) & 3));
        }
        iscf = iscf_buf;
        iscf_st = st;
        iscf_run = iscf_buf;
        /* global gain (minimp3.h:708-709): BITS_DEQUANTIZER_OUT = -1, MAX_SCFI = 44 */
        const int gain_exp = (int)d.global_gain - 4 - 210 - ((d.flags0 & GCF_MS_STEREO) ? 2 : 0);
        gain = l3_ldexp_q2(2048.0f, 44 - gain_exp);

        sfbw = MP3B_ENT_SFB(T, d.n_short_sfb ? (d.n_long_sfb ? 1 : 2) : 0, d.sr_idx);
        sfbw_run = sfbw;
        limit = d.part_23_length;
        mode = d.big_values ? EMODE_BIG : EMODE_C1A;
        big_left = d.big_values;
        band_left = 0;
        ireg = 0;
        for (int i = 0; i < 3; i++) { table_select[i] = d.table_select[i]; region_count[i] = d.region_count[i]; }
        reg_left = region_count[0] + 1;
        tinfo = MP3B_ENT_TABINFO(T, table_select[0]);
        one = 0.f;
        pending = 0; extent = 0;
        c1info = (d.flags0 & GCF_COUNT1_TABLE) ? T.c1info[1] : T.c1info[0];
#if defined(__CUDA_ARCH__)
        asm volatile("" : "+r"(c1info));    /* a register, not four instructions per step rebuilding it from the flags */
#endif
    }

    MP3_HD float band_scale(int b) const { return l3_ldexp_q2(gain, (int)iscf[b*iscf_st] << scf_shift); }

    /* next two lines; q0/q1 receive the signed quantised values.  Big-value pairs and count1 half-quads share
     * one code path (a count1 line is a value of magnitude 0 or 1: one * pow43(1) == one exactly), so lanes in
     * different regions of their granules do not serialise. */
    /* WANT_Q = false (production): the signed integers q0 / q1 are only needed by the parity taps */
    template <bool WANT_Q = true>
    MP3_HD void step(const EntropyTables &T, float &v0, float &v1, int &q0, int &q1)
    {
        v0 = v1 = 0.f;
        q0 = q1 = 0;
        if (mode == EMODE_DONE) return;
        if (band_left == 0)
        {
            /* next scalefactor band (minimp3.h:788-790, 865); region / code-book switch only while in big values */
#if defined(__CUDA_ARCH__)
            /* both run through shared memory on the device; as per-lane pointers in registers they cost one LDS
             * each (an index into a named table would have its base rebuilt at every use) */
            __builtin_assume(__isShared(sfbw_run)); __builtin_assume(__isShared(iscf_run));
#endif
            const int wd = *sfbw_run++;
            if (!wd) { mode = EMODE_DONE; return; }
            band_left = wd >> 1;
            one = l3_ldexp_q2(gain, (int)*iscf_run << scf_shift);
            iscf_run += iscf_st;
            if (mode == EMODE_BIG)
            {
                if (reg_left == 0)
                {
                    ireg++;
                    const int ts = ireg == 1 ? table_select[1] : table_select[2];
                    const int rc = ireg == 1 ? region_count[1] : region_count[2];
                    tinfo = MP3B_ENT_TABINFO(T, ts);
                    reg_left = rc + 1;
                }
                reg_left--;
            }
        }
        /* one look at the stream per step: t0 = the 32 bits at the cursor; the code (19 bits at most) is walked
         * inside t0 with a local count c, and the cursor moves once, at the end */
        br.normalize();
        const uint32_t t0 = br.top();
        int c = 0;
        int x, y, linbits = 0;
        if (mode != EMODE_C1B)
        {
            const bool big = mode == EMODE_BIG;
            const uint32_t info = big ? tinfo : c1info;
            const int base = (int)(info & 0xFFFF), rootw = (int)((info >> 16) & 15);
            int e = 0;
            if (rootw)
            {
                e = MP3B_ENT_LUT(T, base + (t0 >> (32 - rootw)));
                int wcur = rootw;
                while (e & 0x8000)
                {
                    c += wcur;
                    wcur = (e >> 11) & 15;      /* 1..8 */
                    e = MP3B_ENT_LUT(T, base + (e & 0x7FF) + ((t0 << c) >> (32 - wcur)));
                }
                c += e >> 8;
            }
            if (big)
            {
                x = (e >> 4) & 15;
                y = e & 15;
                linbits = (int)((info >> 20) & 15);
            } else
            {
                /* count1 quad: stop when its code ends past the granule's bit budget (minimp3.h:852-864) */
                if (br.bits_consumed() + c > limit) { mode = EMODE_DONE; return; }
                x = (e >> 3) & 1;
                y = (e >> 2) & 1;
                pending = e & 3;
            }
        } else
        {
            x = (pending >> 1) & 1;
            y = pending & 1;
        }
        /* escapes and signs of both values (minimp3.h:807-826) without branches, from the 32 bits behind the
         * code: the reader holds 96 bits from the start of the cursor's word, a step takes 19 + 2*(13 + 1) at most */
        {
            const int n1 = x == 15 ? linbits : 0, s1 = x != 0;
            const int n2 = y == 15 ? linbits : 0, s2 = y != 0;
            const uint32_t t = BitRd::shl_hi(t0, br.top2(), c);
            x += (int)((t >> 1) >> (31 - n1));
            const uint32_t t1 = t << n1;
            const uint32_t neg_x = t1 & (0u - (uint32_t)s1) & 0x80000000u;     /* sign bit in place */
            const uint32_t t2 = t1 << s1;
            y += (int)((t2 >> 1) >> (31 - n2));
            const uint32_t t3 = t2 << n2;
            const uint32_t neg_y = t3 & (0u - (uint32_t)s2) & 0x80000000u;
            br.skip(c + n1 + s1 + n2 + s2);
            /* only a step with escapes can leave the cursor beyond the second word: take one word off now so that
             * the next call's normalize() finds it below 64 */
            if (n1 | n2) br.normalize_wide();
            /* a zero value gives one * 0 = +0 and reads no sign */
            v0 = flip_sign(MP3_FMUL(one, l3_pow43(x, MP3B_ENT_POW43(T))), neg_x);
            v1 = flip_sign(MP3_FMUL(one, l3_pow43(y, MP3B_ENT_POW43(T))), neg_y);
            if (WANT_Q)
            {
                q0 = neg_x ? -x : x;
                q1 = neg_y ? -y : y;
            }
        }
        if (mode == EMODE_BIG) { if (--big_left == 0) mode = EMODE_C1A; }
        else mode = (mode == EMODE_C1A) ? EMODE_C1B : EMODE_C1A;
        band_left--;
        extent += 2;    /* every step before the last one gets here */
    }
};

}  // namespace mp3b
#endif
