/*
 * tests/cpu_harness.cpp -- TEST-ONLY shared object: compiles the product's host parser and the
 * __host__ __device__ arithmetic (l3_math.cuh, l3_entropy.cuh, l3_tables_build.h) for the CPU so that
 * pytest can check, without a GPU, the exact code the CUDA lanes execute against the reference taps.
 * It is not part of libminimp3_b200.so and the product never loads it.
 *
 * h_transform is a plain sequential driver (carried tails / history) around the same building blocks;
 * the tile + halo scheduling of k_transform.cu is checked on the GPU.
 */
#include <stdint.h>
#include <string.h>

#include <vector>

#include "host_parse.h"
#include "l3_entropy.cuh"
#include "l3_math.cuh"
#include "l3_tables_build.h"

using namespace mp3b;

static EntropyTables g_T;
static uint32_t g_tabinfo[32];
static uint16_t g_lut[MP3T_HUFF_LUT_SIZE + 80];
static ImdctConsts g_K;
static uint16_t g_map[16*576];
static float g_fir[32*16];
static bool g_ready = false;

static void ensure_tables()
{
    if (g_ready) return;
    for (int t = 0; t < 32; t++)
        g_tabinfo[t] = (uint32_t)MP3T_HUFF_BASE[t] | ((uint32_t)MP3T_HUFF_ROOTW[t] << 16) | ((uint32_t)MP3T_HUFF_LINBITS[t] << 20);
    for (int i = 0; i < MP3T_HUFF_LUT_SIZE; i++) g_lut[i] = MP3T_HUFF_LUT[i];
    for (int i = 0; i < 64; i++) g_lut[MP3T_HUFF_LUT_SIZE + i] = (uint16_t)(((MP3T_COUNT1A[i] >> 4) << 8) | (MP3T_COUNT1A[i] & 15));
    for (int i = 0; i < 16; i++) g_lut[MP3T_HUFF_LUT_SIZE + 64 + i] = (uint16_t)((4 << 8) | (~i & 15));
    g_T.lut = g_lut; g_T.pow43 = MP3T_POW43; g_T.tabinfo = g_tabinfo;
    g_T.c1info[0] = (uint32_t)MP3T_HUFF_LUT_SIZE | (6u << 16);
    g_T.c1info[1] = (uint32_t)(MP3T_HUFF_LUT_SIZE + 64) | (4u << 16);
    g_T.sfb_long = MP3T_SFB_LONG; g_T.sfb_short = MP3T_SFB_SHORT; g_T.sfb_mixed = MP3T_SFB_MIXED;
    build_consts(&g_K);
    build_reorder_map(g_map);
    build_fir(g_fir);
    g_ready = true;
}

extern "C" {

int h_sizeof_gcdesc() { return (int)sizeof(GcDesc); }
int h_sizeof_framerec() { return (int)sizeof(FrameRec); }
int h_sizeof_tiledesc() { return (int)sizeof(TileDesc); }

/* scalars: [0] ret [1] samples_decoded [2] out_skip [3] out_samples [4] channels [5] hz [6] layer
 *          [7] avg_bitrate [8] main_data_bytes [9] frames [10] frames_decoded [11] n_gc [12] n_tiles
 *          [13] n_granules [14] n_istereo */
int h_walk(const uint8_t *buf, size_t size, int raw, int allow, GcDesc *gcs_out, int gcs_cap, TileDesc *tiles_out,
           int tiles_cap, uint32_t *gr_gc, uint8_t *gr_nch, uint8_t *gr_reset, int gr_cap, uint8_t *md_out,
           int64_t *scalars)
{
    StreamPlan plan;
    WalkOptions opt;
    opt.raw_frames = raw != 0;
    opt.allow_mono_stereo = allow != 0;
    std::vector<uint8_t> md(size + 64, 0);
    walk_stream(buf, size, md.data(), 0, opt, &plan);
    scalars[0] = plan.ret; scalars[1] = (int64_t)plan.samples_decoded; scalars[2] = (int64_t)plan.out_skip;
    scalars[3] = (int64_t)plan.out_samples; scalars[4] = plan.channels; scalars[5] = plan.hz; scalars[6] = plan.layer;
    scalars[7] = plan.avg_bitrate_kbps; scalars[8] = (int64_t)plan.main_data_bytes; scalars[9] = plan.frames;
    scalars[10] = plan.frames_decoded; scalars[11] = (int64_t)plan.gcs.size(); scalars[12] = (int64_t)plan.tiles.size();
    scalars[13] = (int64_t)plan.granules.size(); scalars[14] = (int64_t)plan.istereo_gcs.size();
    if (gcs_out) for (size_t i = 0; i < plan.gcs.size() && (int)i < gcs_cap; i++) gcs_out[i] = plan.gcs[i];
    if (tiles_out) for (size_t i = 0; i < plan.tiles.size() && (int)i < tiles_cap; i++) tiles_out[i] = plan.tiles[i];
    if (gr_gc)
        for (size_t i = 0; i < plan.granules.size() && (int)i < gr_cap; i++)
        {
            gr_gc[i] = plan.granules[i].gc; gr_nch[i] = plan.granules[i].nch; gr_reset[i] = plan.granules[i].reset;
        }
    if (md_out) memcpy(md_out, md.data(), size + 16);
    return 0;
}

/* The walker in device_side_info mode + the GPU's descriptor builder (frame_descs, l3_sideinfo.cuh) run on the host:
 * gcs_out receives what k_l3_side_info would write, order_out the lane order.  scalars as h_walk (n_gc in [11]). */
int h_walk_frame_recs(const uint8_t *buf, size_t size, int raw, int allow, GcDesc *gcs_out, int gcs_cap, uint32_t *order_out,
                      int order_cap, int64_t *scalars)
{
    StreamPlan plan;
    WalkOptions opt;
    opt.raw_frames = raw != 0;
    opt.allow_mono_stereo = allow != 0;
    opt.device_side_info = true;
    std::vector<uint8_t> md(size + 64, 0);
    walk_stream(buf, size, md.data(), 0, opt, &plan);
    scalars[0] = plan.ret; scalars[1] = (int64_t)plan.samples_decoded; scalars[2] = (int64_t)plan.out_skip;
    scalars[3] = (int64_t)plan.out_samples; scalars[4] = plan.channels; scalars[5] = plan.hz; scalars[6] = plan.layer;
    scalars[7] = plan.avg_bitrate_kbps; scalars[8] = (int64_t)plan.main_data_bytes; scalars[9] = plan.frames;
    scalars[10] = plan.frames_decoded; scalars[11] = (int64_t)plan.n_gc; scalars[12] = (int64_t)plan.tiles.size();
    scalars[13] = (int64_t)plan.granules.size(); scalars[14] = (int64_t)plan.istereo_gcs.size();
    if (!plan.gcs.empty() || plan.gc_key.size() != plan.n_gc) return -1;
    for (const FrameRec &fr : plan.frame_recs)
    {
        GcDesc d[4];
        const int n = frame_descs(fr, d);
        if (n != fr.n_gr*fr.nch) return -2;
        for (int k = 0; k < n; k++)
            if ((int)(fr.gc_first + k) < gcs_cap) gcs_out[fr.gc_first + k] = d[k];
        if (fr.pad[0] && (int)fr.gc_first - 1 < gcs_cap)
        {
            GcDesc pad;
            memset(&pad, 0, sizeof(pad));
            pad.flags1 = GCF1_MPEG1;
            gcs_out[fr.gc_first - 1] = pad;
        }
    }
    if (order_out) for (size_t i = 0; i < plan.order.size() && (int)i < order_cap; i++) order_out[i] = plan.order[i];
    return (int)plan.order.size();
}

/* the per-lane entropy decode, one granule-channel after the other */
void h_entropy(const uint8_t *md, size_t md_bytes, const GcDesc *gcs, int n_gc, int16_t *q, float *xr, float *scf,
               uint8_t *ist, int *extent)
{
    ensure_tables();
    std::vector<uint32_t> words((md_bytes + 3)/4 + 2, 0);
    memcpy(words.data(), md, md_bytes);
    for (int g = 0; g < n_gc; g++)
    {
        const GcDesc &d = gcs[g];
        uint8_t iscf[40], prev[40], istl[40];
        memset(iscf, 0, sizeof(iscf)); memset(prev, 0, sizeof(prev)); memset(istl, 0, sizeof(istl));
        int n_ist = 0;
        const bool use_d0 = (d.flags1 & GCF1_GR1) && d.scfsi;
        const GcDesc *d0 = use_d0 ? &gcs[g - ((d.flags1 & GCF1_STEREO) ? 2 : 1)] : nullptr;
        GranuleDecoder dec;
        dec.init(d, d0, words.data(), words.size(), g_T, iscf, prev, 1, istl, &n_ist);
        const int n_sfb = d.n_long_sfb + d.n_short_sfb;
        for (int i = 0; i < 40; i++)
        {
            scf[(size_t)g*40 + i] = i < n_sfb ? dec.band_scale(i) : 0.f;
            ist[(size_t)g*40 + i] = i < n_ist ? istl[i] : 0;
        }
        for (int it = 0; it < 288; it++)
        {
            float v0, v1;
            int q0, q1;
            dec.step(g_T, v0, v1, q0, q1);
            xr[(size_t)g*576 + 2*it] = v0; xr[(size_t)g*576 + 2*it + 1] = v1;
            q[(size_t)g*576 + 2*it] = (int16_t)q0; q[(size_t)g*576 + 2*it + 1] = (int16_t)q1;
        }
        extent[g] = dec.extent;
    }
}

/* reorder + antialias + IMDCT + sign + DCT-II + window over a granule list, carried state (sequential) */
void h_transform(const float *xr_stereo, const GcDesc *gcs, const uint32_t *gr_gc, const uint8_t *gr_nch,
                 const uint8_t *gr_reset, int n_gr, float *xr_aa, float *sub, int16_t *pcm16, float *pcmf)
{
    ensure_tables();
    float tails[2][32][9];
    static float hist[2][15][32];
    memset(tails, 0, sizeof(tails));
    memset(hist, 0, sizeof(hist));
    size_t out = 0;
    for (int g = 0; g < n_gr; g++)
    {
        const int nch = gr_nch[g];
        if (gr_reset[g]) { memset(tails, 0, sizeof(tails)); memset(hist, 0, sizeof(hist)); }
        float y[2][18][32];
        for (int ch = 0; ch < nch; ch++)
        {
            const uint32_t gc = gr_gc[g] + ch;
            const GcDesc &d = gcs[gc];
            const int bt = d.flags0 & GCF_BLOCK_TYPE_MASK;
            const bool mixed = d.flags0 & GCF_MIXED;
            const int n_long = mixed ? (d.sr_idx == 1 ? 4 : 2) : 0;
            float x[576];
            const float *src = xr_stereo + (size_t)gc*576;
            if (bt == 2)
            {
                const uint16_t *map = g_map + (d.sr_idx*2 + (mixed ? 1 : 0))*576;
                for (int e = 0; e < 576; e++) x[map[e]] = src[e];
            } else memcpy(x, src, sizeof(x));
            const int aa_bands = bt == 2 ? n_long - 1 : 31;
            for (int b = 0; b < aa_bands; b++)
                for (int i = 0; i < 8; i++)
                {
                    const float u = x[(b + 1)*18 + i], dn = x[b*18 + 17 - i];
                    x[(b + 1)*18 + i] = u*g_K.aa_cs[i] - dn*g_K.aa_ca[i];
                    x[b*18 + 17 - i] = u*g_K.aa_ca[i] + dn*g_K.aa_cs[i];
                }
            if (xr_aa) memcpy(xr_aa + (size_t)gc*576, x, sizeof(x));
            for (int sb = 0; sb < 32; sb++)
            {
                float own[9], tail[9], o[18];
                const bool is_short = bt == 2 && sb >= n_long;
                if (is_short) imdct_short_core(x + sb*18, own, tail, g_K.twid3);
                else imdct36_core(x + sb*18, own, tail, g_K.twid9);
                if (is_short) imdct_short_combine(tails[ch][sb], own, g_K.twid3, o);
                else imdct36_combine(tails[ch][sb], own, g_K.win[(bt == 3 && sb >= n_long) ? 1 : 0], o);
                memcpy(tails[ch][sb], tail, sizeof(tail));
                if (sb & 1) for (int k = 1; k < 18; k += 2) o[k] = -o[k];
                for (int k = 0; k < 18; k++) y[ch][k][sb] = o[k];
                if (sub) for (int k = 0; k < 18; k++) sub[(size_t)gc*576 + sb*18 + k] = o[k];
            }
            for (int t = 0; t < 18; t++) dct2_32(y[ch][t], g_K.dct_sec);
        }
        for (int ch = 0; ch < nch; ch++)
        {
            /* rows: 15 history + 18 new */
            float rows[33][32];
            memcpy(rows, hist[ch], sizeof(hist[ch]));
            memcpy(rows + 15, y[ch], sizeof(y[ch]));
            for (int t = 0; t < 18; t++)
                for (int j = 0; j < 32; j++)
                {
                    const int ia = j < 16 ? 16 + j : (j == 16 ? 0 : 48 - j);
                    const int ib = j <= 16 ? 16 - j : j - 16;
                    float acc = 0.f;
                    for (int m = 0; m < 8; m++)
                    {
                        acc = fmaf(g_fir[j*16 + 2*m], rows[15 + t - 2*m][ia], acc);
                        acc = fmaf(g_fir[j*16 + 2*m + 1], rows[15 + t - 2*m - 1][ib], acc);
                    }
                    const size_t idx = out + (size_t)(t*32 + j)*nch + ch;
                    if (pcm16) pcm16[idx] = (j == 0 || j == 16) ? pcm_round_scalar(acc) : pcm_round_simd(acc);
                    if (pcmf) pcmf[idx] = acc*(1.0f/32768.0f);
                }
            memcpy(hist[ch], rows + 18, sizeof(hist[ch]));
        }
        out += (size_t)576*nch;
    }
}

float h_pow43(int x) { ensure_tables(); return l3_pow43(x, MP3T_POW43); }
float h_ldexp_q2(float y, int e) { return l3_ldexp_q2(y, e); }
void h_dct32(float *x) { ensure_tables(); dct2_32(x, g_K.dct_sec); }
int h_detect_buf(const uint8_t *buf, size_t size) { return detect_buf(buf, size); }

}  // extern "C"
