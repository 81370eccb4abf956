/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin C-ABI shim around the UNMODIFIED reference decoder.  The reference
 * headers are #included from where they lie (-I/root/reference); no reference
 * source is copied into this repository.  Built by oracle/Makefile into
 * oracle/_ref/libminimp3_ref_{s16,f32}.so.
 *
 * What it exports (used by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg only):
 *   ref_decode_frames   raw mp3dec_decode_frame loop (minimp3.h:1713-1806)
 *   ref_load_buf        mp3dec_load_buf               (minimp3_ex.h:308-525)
 *   ref_tap_stream      stage taps: drives the reference's own static stage
 *                       functions (L3_read_side_info, L3_decode_scalefactors,
 *                       L3_huffman, stereo, reorder, antialias, IMDCT, synth)
 *                       in the order mp3dec_decode_frame/L3_decode use them,
 *                       copying out every intermediate buffer, and cross-checks
 *                       its PCM against the real mp3dec_decode_frame.
 *   ref_bench_decode    multi-threaded CPU baseline (one mp3dec_t per stream,
 *                       plain mp3dec_init + mp3dec_decode_frame loop)
 *   ref_const_*         dumps of the reference's constant tables / helper
 *                       functions, used to verify the closed-form table
 *                       generator (tools/gen_tables.py)
 */
#define MINIMP3_IMPLEMENTATION
#define MINIMP3_ALLOW_MONO_STEREO_TRANSITION
#include "minimp3_ex.h"

#include <pthread.h>
#include <stdio.h>
#include <time.h>

typedef struct
{
    int offset;        /* absolute byte offset of the call's input pointer */
    int frame_bytes;   /* info.frame_bytes (bytes consumed by the call)     */
    int frame_offset;  /* info.frame_offset (garbage skipped before frame)  */
    int samples;       /* return value (per channel)                         */
    int channels, hz, layer, bitrate_kbps;
} ref_frame_rec;

int ref_sample_size(void) { return (int)sizeof(mp3d_sample_t); }

/* Raw frame loop, as in the README usage and minimp3_ex.h:445-511 (buffer mode). */
long ref_decode_frames(const uint8_t *buf, long size, mp3d_sample_t *pcm, long pcm_cap,
                       ref_frame_rec *frames, long frames_cap, long *n_frames_out)
{
    mp3dec_t dec;
    mp3dec_frame_info_t info;
    mp3d_sample_t tmp[MINIMP3_MAX_SAMPLES_PER_FRAME];
    long pos = 0, total = 0, nf = 0;
    mp3dec_init(&dec);
    memset(&info, 0, sizeof(info));
    do
    {
        int n = (int)MINIMP3_MIN(size - pos, (long)INT_MAX);
        int samples = mp3dec_decode_frame(&dec, buf + pos, n, tmp, &info);
        if (frames && nf < frames_cap)
        {
            ref_frame_rec *r = frames + nf;
            r->offset = (int)pos; r->frame_bytes = info.frame_bytes; r->frame_offset = info.frame_offset;
            r->samples = samples; r->channels = info.channels; r->hz = info.hz; r->layer = info.layer;
            r->bitrate_kbps = info.bitrate_kbps;
        }
        nf++;
        if (samples)
        {
            long n_out = (long)samples*info.channels;
            if (pcm && total + n_out <= pcm_cap)
                memcpy(pcm + total, tmp, n_out*sizeof(mp3d_sample_t));
            total += n_out;
        }
        pos += info.frame_bytes;
    } while (info.frame_bytes);
    if (n_frames_out) *n_frames_out = nf;
    return total;
}

int ref_load_buf(const uint8_t *buf, size_t size, mp3d_sample_t **out, size_t *samples,
                 int *channels, int *hz, int *layer, int *avg_bitrate_kbps)
{
    mp3dec_t dec;
    mp3dec_file_info_t info;
    int ret;
    memset(&info, 0, sizeof(info));
    ret = mp3dec_load_buf(&dec, buf, size, &info, 0, 0);
    *out = info.buffer; *samples = info.samples; *channels = info.channels; *hz = info.hz;
    *layer = info.layer; *avg_bitrate_kbps = info.avg_bitrate_kbps;
    return ret;
}

void ref_free(void *p) { free(p); }

int ref_detect_buf(const uint8_t *buf, size_t size) { return mp3dec_detect_buf(buf, size); }

/* ------------------------------------------------------------------ taps */

typedef struct
{
    int frame, gr, ch, nch;
    int hdr;                 /* 4 header bytes, big endian */
    int bit_pos;             /* bs.pos in scratch.maindata at granule-channel start */
    int main_data_begin;
    int part_23_length, big_values, scalefac_compress;
    int global_gain, block_type, mixed_block_flag, n_long_sfb, n_short_sfb;
    int table_select[3], region_count[3], subblock_gain[3];
    int preflag, scalefac_scale, count1_table, scfsi;
    int sfb_kind;            /* 0 long, 1 short, 2 mixed */
    int sr_idx;              /* row of g_scf_* used (0..7) */
    int pcm_offset;          /* sample offset (channel-inclusive) of this granule in the output */
} ref_gc_rec;

typedef struct
{
    long cap;            /* capacity in granule-channels */
    ref_gc_rec *rec;     /* [cap]            */
    float *scf;          /* [cap][40]        L3_decode_scalefactors output     */
    uint8_t *ist_pos;    /* [cap][40]        (39 used)                          */
    int16_t *q;          /* [cap][576]       signed quantised lines            */
    float *xr;           /* [cap][576]       after L3_huffman (requantised)    */
    float *xr_stereo;    /* [cap][576]       after MS / intensity stereo       */
    float *xr_aa;        /* [cap][576]       after reorder + antialias         */
    float *sub;          /* [cap][576]       after IMDCT + overlap + sign      */
} ref_taps;

static float g_p43_inv_tab[8207];
static int g_p43_inv_ready;

static int ref_invert_pow43(float a)
{
    int lo = 0, hi = 8206;
    if (!g_p43_inv_ready)
    {
        int i;
        for (i = 0; i <= 8206; i++) g_p43_inv_tab[i] = L3_pow_43(i);
        g_p43_inv_ready = 1;
    }
    while (lo < hi)
    {
        int mid = (lo + hi)/2;
        if (g_p43_inv_tab[mid] < a) lo = mid + 1; else hi = mid;
    }
    return (g_p43_inv_tab[lo] == a) ? lo : -1;
}

/*
 * Mirrors mp3dec_decode_frame (minimp3.h:1713-1806) for Layer III using the
 * reference's own static functions, with buffer captures between stages.
 * Returns number of granule-channels captured, or <0 on internal mismatch
 * against the real mp3dec_decode_frame (-2: sample count, -3: pcm, -4: q inversion).
 */
long ref_tap_stream(const uint8_t *buf, long size, ref_taps *t, mp3d_sample_t *pcm_out, long pcm_cap, long *total_samples_out)
{
    mp3dec_t dec, chk;
    mp3dec_frame_info_t info, info2;
    static mp3dec_scratch_t scratch;   /* zero-initialised like nothing in the reference; see note below */
    mp3d_sample_t pcm_a[MINIMP3_MAX_SAMPLES_PER_FRAME], pcm_b[MINIMP3_MAX_SAMPLES_PER_FRAME];
    long pos = 0, total = 0, ngc = 0;
    int frame_no = 0;
    mp3dec_init(&dec);
    mp3dec_init(&chk);
    memset(&info, 0, sizeof(info));
    memset(&scratch, 0, sizeof(scratch));
    do
    {
        const uint8_t *mp3 = buf + pos;
        int mp3_bytes = (int)MINIMP3_MIN(size - pos, (long)INT_MAX);
        int i = 0, igr, frame_size = 0, success = 1, samples = 0, samples_chk;
        const uint8_t *hdr;
        bs_t bs_frame[1];

        samples_chk = mp3dec_decode_frame(&chk, mp3, mp3_bytes, pcm_b, &info2);

        /* --- sync, exactly the steps of minimp3.h:1720-1746 --- */
        if (mp3_bytes > 4 && dec.header[0] == 0xff && hdr_compare(dec.header, mp3))
        {
            frame_size = hdr_frame_bytes(mp3, dec.free_format_bytes) + hdr_padding(mp3);
            if (frame_size != mp3_bytes && (frame_size + HDR_SIZE > mp3_bytes || !hdr_compare(mp3, mp3 + frame_size)))
                frame_size = 0;
        }
        if (!frame_size)
        {
            memset(&dec, 0, sizeof(mp3dec_t));
            i = mp3d_find_frame(mp3, mp3_bytes, &dec.free_format_bytes, &frame_size);
            if (!frame_size || i + frame_size > mp3_bytes)
            {
                info.frame_bytes = i;
                goto frame_done;
            }
        }
        hdr = mp3 + i;
        memcpy(dec.header, hdr, HDR_SIZE);
        info.frame_bytes = i + frame_size;
        info.frame_offset = i;
        info.channels = HDR_IS_MONO(hdr) ? 1 : 2;
        info.hz = hdr_sample_rate_hz(hdr);
        info.layer = 4 - HDR_GET_LAYER(hdr);
        info.bitrate_kbps = hdr_bitrate_kbps(hdr);

        bs_init(bs_frame, hdr + HDR_SIZE, frame_size - HDR_SIZE);
        if (HDR_IS_CRC(hdr))
            get_bits(bs_frame, 16);

        if (info.layer == 3)
        {
            mp3d_sample_t *pcm = pcm_a;
            int main_data_begin = L3_read_side_info(bs_frame, scratch.gr_info, hdr);
            if (main_data_begin < 0 || bs_frame->pos > bs_frame->limit)
            {
                mp3dec_init(&dec);
                goto frame_done;
            }
            success = L3_restore_reservoir(&dec, bs_frame, &scratch, main_data_begin);
            if (success)
            {
                int nch = info.channels;
                for (igr = 0; igr < (HDR_TEST_MPEG1(hdr) ? 2 : 1); igr++, pcm += 576*nch)
                {
                    L3_gr_info_t *gr_info = scratch.gr_info + igr*nch;
                    int ch;
                    long base = ngc;
                    memset(scratch.grbuf[0], 0, 576*2*sizeof(float));
                    /* L3_decode (minimp3.h:1238-1272), stage by stage */
                    for (ch = 0; ch < nch; ch++)
                    {
                        int layer3gr_limit = scratch.bs.pos + gr_info[ch].part_23_length;
                        bs_t bs_save = scratch.bs;
                        long k = ngc + ch;
                        int sr_idx = HDR_GET_MY_SAMPLE_RATE(hdr); sr_idx -= (sr_idx != 0);
                        L3_decode_scalefactors(dec.header, scratch.ist_pos[ch], &scratch.bs, gr_info + ch, scratch.scf, ch);
                        if (t && k < t->cap)
                        {
                            ref_gc_rec *r = t->rec + k;
                            const L3_gr_info_t *g = gr_info + ch;
                            int j;
                            bs_t bs_q = scratch.bs;
                            float ones[40], tmp[576 + 4];
                            r->frame = frame_no; r->gr = igr; r->ch = ch; r->nch = nch;
                            r->hdr = (hdr[0] << 24) | (hdr[1] << 16) | (hdr[2] << 8) | hdr[3];
                            r->bit_pos = bs_save.pos; r->main_data_begin = main_data_begin;
                            r->part_23_length = g->part_23_length; r->big_values = g->big_values;
                            r->scalefac_compress = g->scalefac_compress; r->global_gain = g->global_gain;
                            r->block_type = g->block_type; r->mixed_block_flag = g->mixed_block_flag;
                            r->n_long_sfb = g->n_long_sfb; r->n_short_sfb = g->n_short_sfb;
                            for (j = 0; j < 3; j++)
                            {
                                r->table_select[j] = g->table_select[j]; r->region_count[j] = g->region_count[j];
                                r->subblock_gain[j] = g->subblock_gain[j];
                            }
                            r->preflag = g->preflag; r->scalefac_scale = g->scalefac_scale;
                            r->count1_table = g->count1_table; r->scfsi = g->scfsi;
                            r->sfb_kind = !g->n_short_sfb ? 0 : (g->n_long_sfb ? 2 : 1);
                            r->sr_idx = sr_idx;
                            r->pcm_offset = (int)(total + (long)igr*576*nch);
                            memcpy(t->scf + k*40, scratch.scf, 40*sizeof(float));
                            memcpy(t->ist_pos + k*40, scratch.ist_pos[ch], 39);
                            /* integer spectrum: same bits decoded with unit scalefactors, inverted through pow43 */
                            for (j = 0; j < 40; j++) ones[j] = 1.0f;
                            memset(tmp, 0, sizeof(tmp));
                            L3_huffman(tmp, &bs_q, g, ones, layer3gr_limit);
                            for (j = 0; j < 576; j++)
                            {
                                float a = tmp[j];
                                int m = ref_invert_pow43(a < 0 ? -a : a);
                                if (m < 0) return -4;
                                t->q[k*576 + j] = (int16_t)(a < 0 ? -m : m);
                            }
                        }
                        L3_huffman(scratch.grbuf[ch], &scratch.bs, gr_info + ch, scratch.scf, layer3gr_limit);
                        if (t && k < t->cap)
                            memcpy(t->xr + k*576, scratch.grbuf[ch], 576*sizeof(float));
                    }
                    if (HDR_TEST_I_STEREO(dec.header))
                        L3_intensity_stereo(scratch.grbuf[0], scratch.ist_pos[1], gr_info, dec.header);
                    else if (HDR_IS_MS_STEREO(dec.header))
                        L3_midside_stereo(scratch.grbuf[0], 576);
                    for (ch = 0; ch < nch; ch++)
                        if (t && base + ch < t->cap)
                            memcpy(t->xr_stereo + (base + ch)*576, scratch.grbuf[ch], 576*sizeof(float));
                    for (ch = 0; ch < nch; ch++)
                    {
                        const L3_gr_info_t *g = gr_info + ch;
                        int aa_bands = 31;
                        int n_long_bands = (g->mixed_block_flag ? 2 : 0) << (int)(HDR_GET_MY_SAMPLE_RATE(dec.header) == 2);
                        if (g->n_short_sfb)
                        {
                            aa_bands = n_long_bands - 1;
                            L3_reorder(scratch.grbuf[ch] + n_long_bands*18, scratch.syn[0], g->sfbtab + g->n_long_sfb);
                        }
                        L3_antialias(scratch.grbuf[ch], aa_bands);
                        if (t && base + ch < t->cap)
                            memcpy(t->xr_aa + (base + ch)*576, scratch.grbuf[ch], 576*sizeof(float));
                        L3_imdct_gr(scratch.grbuf[ch], dec.mdct_overlap[ch], g->block_type, n_long_bands);
                        L3_change_sign(scratch.grbuf[ch]);
                        if (t && base + ch < t->cap)
                            memcpy(t->sub + (base + ch)*576, scratch.grbuf[ch], 576*sizeof(float));
                    }
                    ngc += nch;
                    mp3d_synth_granule(dec.qmf_state, scratch.grbuf[0], 18, nch, pcm, scratch.syn[0]);
                }
            }
            L3_save_reservoir(&dec, &scratch);
            samples = success*hdr_frame_samples(dec.header);
        } else
        {
            /* Layer I/II: no taps, keep decoder states in lock-step via the real entry point */
            dec = chk;
            samples = samples_chk;
            memcpy(pcm_a, pcm_b, sizeof(pcm_a));
            info = info2;
        }
frame_done:
        if (samples != samples_chk || info.frame_bytes != info2.frame_bytes)
            return -2;
        if (samples)
        {
            long n_out = (long)samples*info.channels;
            if (memcmp(pcm_a, pcm_b, n_out*sizeof(mp3d_sample_t)))
                return -3;
            if (pcm_out && total + n_out <= pcm_cap)
                memcpy(pcm_out + total, pcm_a, n_out*sizeof(mp3d_sample_t));
            total += n_out;
        }
        pos += info.frame_bytes;
        frame_no++;
    } while (info.frame_bytes);
    if (total_samples_out) *total_samples_out = total;
    return ngc;
}

/* ------------------------------------------------------- synthesis-only tap */

/* Runs mp3d_synth_granule (minimp3.h:1629-1655) over a sequence of n_gr granules of
 * subband samples sub[n_gr][nch][576] starting from a zero qmf_state. */
void ref_synth_sequence(const float *sub, int n_gr, int nch, mp3d_sample_t *pcm)
{
    float qmf_state[15*2*32];
    float grbuf[2][576];
    float syn[18 + 15][2*32];
    int g;
    memset(qmf_state, 0, sizeof(qmf_state));
    for (g = 0; g < n_gr; g++)
    {
        memset(grbuf, 0, sizeof(grbuf));
        memcpy(grbuf[0], sub + (size_t)g*nch*576, sizeof(float)*576*nch);
        mp3d_synth_granule(qmf_state, grbuf[0], 18, nch, pcm + (size_t)g*576*nch, syn[0]);
    }
}

/* Runs antialias + IMDCT + sign over a sequence of granules of one channel from zero overlap. */
void ref_imdct_sequence(const float *xr_aa, int n_gr, const int *block_type, const int *n_long_bands, float *sub)
{
    float overlap[9*32];
    float grbuf[576];
    int g;
    memset(overlap, 0, sizeof(overlap));
    for (g = 0; g < n_gr; g++)
    {
        memcpy(grbuf, xr_aa + (size_t)g*576, sizeof(grbuf));
        L3_imdct_gr(grbuf, overlap, block_type[g], n_long_bands[g]);
        L3_change_sign(grbuf);
        memcpy(sub + (size_t)g*576, grbuf, sizeof(grbuf));
    }
}

void ref_dct_ii(float *grbuf, int n) { mp3d_DCT_II(grbuf, n); }

/* ------------------------------------------------------------ constants */

float ref_const_pow43(int x) { return L3_pow_43(x); }
float ref_const_ldexp_q2(float y, int e) { return L3_ldexp_q2(y, e); }

/* sfb width tables, via a fabricated side info (kind: 0 long 1 short 2 mixed); returns count incl. terminator */
int ref_const_sfb_table(int hdr_be, int kind, uint8_t *out, int cap)
{
    /* build a minimal side-info bitstream that selects the wanted table */
    uint8_t hdr[4], si[40];
    bs_t bs[1];
    L3_gr_info_t gr[4];
    int n = 0, mpeg1, mono, bitpos;
    hdr[0] = (uint8_t)(hdr_be >> 24); hdr[1] = (uint8_t)(hdr_be >> 16); hdr[2] = (uint8_t)(hdr_be >> 8); hdr[3] = (uint8_t)hdr_be;
    mpeg1 = HDR_TEST_MPEG1(hdr) != 0; mono = HDR_IS_MONO(hdr);
    memset(si, 0, sizeof(si));
    /* position of window_switching_flag of the first granule-channel */
    bitpos = mpeg1 ? (9 + (mono ? 5 : 3) + (mono ? 4 : 8)) : (8 + (mono ? 1 : 2));
    bitpos += 12 + 9 + 8 + (mpeg1 ? 4 : 9);
    if (kind)
    {
        /* window_switching=1, block_type=2, mixed=(kind==2) */
        int bits = (1 << 3) | (2 << 1) | (kind == 2);
        int k;
        for (k = 0; k < 4; k++)
            if (bits & (8 >> k)) si[(bitpos + k) >> 3] |= (uint8_t)(0x80 >> ((bitpos + k) & 7));
    }
    bs_init(bs, si, sizeof(si));
    if (L3_read_side_info(bs, gr, hdr) < 0) return -1;
    while (n < cap) { out[n] = gr[0].sfbtab[n]; if (!out[n++]) break; }
    return n;
}

/* ------------------------------------------------------------ CPU baseline */

typedef struct
{
    const uint8_t *const *bufs;
    const long *sizes;
    int first, last;
    long samples;
} bench_job;

static void *bench_worker(void *arg)
{
    bench_job *j = (bench_job *)arg;
    mp3d_sample_t pcm[MINIMP3_MAX_SAMPLES_PER_FRAME];
    int s;
    long total = 0;
    for (s = j->first; s < j->last; s++)
    {
        mp3dec_t dec;
        mp3dec_frame_info_t info;
        const uint8_t *p = j->bufs[s];
        long left = j->sizes[s];
        mp3dec_init(&dec);
        do
        {
            int samples = mp3dec_decode_frame(&dec, p, (int)MINIMP3_MIN(left, (long)INT_MAX), pcm, &info);
            total += (long)samples*info.channels;
            p += info.frame_bytes; left -= info.frame_bytes;
        } while (info.frame_bytes);
    }
    j->samples = total;
    return 0;
}

/* Decodes n streams with `threads` pthreads (contiguous stream ranges); returns
 * channel-inclusive samples decoded; *seconds = wall-clock of the parallel region. */
long ref_bench_decode(const uint8_t *const *bufs, const long *sizes, int n, int threads, double *seconds)
{
    pthread_t *th;
    bench_job *jobs;
    struct timespec t0, t1;
    int i;
    long total = 0;
    if (threads < 1) threads = 1;
    if (threads > n) threads = n > 0 ? n : 1;
    th = (pthread_t *)malloc(sizeof(pthread_t)*threads);
    jobs = (bench_job *)malloc(sizeof(bench_job)*threads);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (i = 0; i < threads; i++)
    {
        jobs[i].bufs = bufs; jobs[i].sizes = sizes;
        jobs[i].first = (int)((long)n*i/threads); jobs[i].last = (int)((long)n*(i + 1)/threads);
        jobs[i].samples = 0;
        pthread_create(&th[i], 0, bench_worker, &jobs[i]);
    }
    for (i = 0; i < threads; i++)
    {
        pthread_join(th[i], 0);
        total += jobs[i].samples;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9*(double)(t1.tv_nsec - t0.tv_nsec);
    free(th); free(jobs);
    return total;
}

/* ------------------------------------------------------------ streaming / seek API (minimp3_ex.h:700-1013) */

/* open_buf + optional seek + read in portions, like minimp3_test.c modes 6-8.  Returns samples read. */
long ref_ex_decode(const uint8_t *buf, size_t size, int flags, uint64_t seek_pos, size_t portion, mp3d_sample_t *out,
                   long cap, uint64_t *total_samples, int *open_ret, int *last_error, int *channels, int *hz)
{
    static mp3dec_ex_t dec;
    long got = 0;
    *open_ret = mp3dec_ex_open_buf(&dec, buf, size, flags);
    *total_samples = 0;
    *last_error = 0;
    if (*open_ret) return 0;
    *total_samples = dec.samples;
    *channels = dec.info.channels;
    *hz = dec.info.hz;
    if (seek_pos && mp3dec_ex_seek(&dec, seek_pos)) { *last_error = -100; mp3dec_ex_close(&dec); return 0; }
    if (!portion) portion = 1152;
    while (got + (long)portion <= cap)
    {
        size_t n = mp3dec_ex_read(&dec, out + got, portion);
        got += (long)n;
        if (n != portion) break;
    }
    *last_error = dec.last_error;
    mp3dec_ex_close(&dec);
    return got;
}

typedef struct { int n; uint64_t offsets[4096]; int frame_sizes[4096]; } ref_iter_log;
static int ref_iter_cb(void *user, const uint8_t *frame, int frame_size, int free_format_bytes, size_t buf_size, uint64_t offset, mp3dec_frame_info_t *info)
{
    ref_iter_log *l = (ref_iter_log *)user;
    (void)frame; (void)free_format_bytes; (void)buf_size; (void)info;
    if (l->n < 4096) { l->offsets[l->n] = offset; l->frame_sizes[l->n] = frame_size; }
    l->n++;
    return 0;
}
int ref_iterate(const uint8_t *buf, size_t size, uint64_t *offsets, int *frame_sizes, int cap)
{
    static ref_iter_log l;
    int i;
    l.n = 0;
    mp3dec_iterate_buf(buf, size, ref_iter_cb, &l);
    for (i = 0; i < l.n && i < cap && i < 4096; i++) { offsets[i] = l.offsets[i]; frame_sizes[i] = l.frame_sizes[i]; }
    return l.n;
}
