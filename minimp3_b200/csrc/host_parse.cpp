/*
 * host_parse.cpp -- see host_parse.h.  Integer-only host logic; every function names the reference
 * lines whose behaviour it restates (paths relative to /root/reference).
 */
#include "host_parse.h"

#include <limits.h>
#include <string.h>

#include <algorithm>

#include "l12_decode.cuh"
#include "mp3_hdr.cuh"
#include "mp3_tables.h"

namespace mp3b {

static const int kHdrSize = 4;
static const int kMaxFreeFormatFrame = 2304;   /* minimp3.h:49 */
static const int kMaxSyncMatches = 10;         /* minimp3.h:50-52 */
static const int kMaxReservoir = 511;          /* minimp3.h:56 */

/* ---- header math: shared with the GPU frame scan (mp3_hdr.cuh) ------------------------------------------------ */

bool hdr_valid(const uint8_t *h) { return hd_valid(h); }
bool hdr_compare(const uint8_t *h1, const uint8_t *h2) { return hd_compare(h1, h2); }
unsigned hdr_bitrate_kbps(const uint8_t *h) { return hd_bitrate_kbps(h); }
unsigned hdr_sample_rate_hz(const uint8_t *h) { return hd_sample_rate_hz(h); }
unsigned hdr_frame_samples(const uint8_t *h) { return hd_frame_samples(h); }
int hdr_frame_bytes(const uint8_t *h, int free_format_size) { return hd_frame_bytes(h, free_format_size); }
int hdr_padding(const uint8_t *h) { return hd_padding(h); }

/* ---- frame sync (behaviour of mp3d_find_frame / mp3d_match_frame, minimp3.h:1657-1706) -------------------------
 * A position is accepted as a frame start when its header is valid and the frames it implies line up: up to
 * kMaxSyncMatches following headers, each found exactly one frame length further, agree with it in version, layer,
 * sample rate and free-format-ness (running out of input after at least one of them is fine); or when the frame
 * fills the input exactly from position 0.  Free-format frames carry no length: it is taken from the distance to
 * the next agreeing header, provided a third one follows at the same distance. */

/* bytes from this header to the next one */
static inline int stride_of(const uint8_t *h, int free_format_bytes)
{
    return hdr_frame_bytes(h, free_format_bytes) + hdr_padding(h);
}

/* avail = bytes from hdr to the end of the input */
static bool followers_agree(const uint8_t *hdr, int avail, int free_format_bytes)
{
    int at = 0;
    for (int seen = 0; seen < kMaxSyncMatches; seen++)
    {
        at += stride_of(hdr + at, free_format_bytes);
        if (at + kHdrSize > avail) return seen != 0;
        if (!hdr_compare(hdr, hdr + at)) return false;
    }
    return true;
}

/* free-format length search: returns the stride (0: none found) and the frame's own byte count without padding.
 * The candidates keep at least another stride plus a header behind them (the reference's loop bound). */
static int free_format_stride(const uint8_t *hdr, int avail, int *frame_bytes)
{
    for (int stride = kHdrSize; stride < kMaxFreeFormatFrame && 2*stride < avail - kHdrSize; stride++)
    {
        if (!hdr_compare(hdr, hdr + stride)) continue;
        const int own = stride - hdr_padding(hdr);                 /* this frame without its padding byte */
        const int next = own + hdr_padding(hdr + stride);          /* the stride of the frame behind it */
        if (stride + next + kHdrSize > avail || !hdr_compare(hdr, hdr + stride + next)) continue;
        *frame_bytes = own;
        return stride;
    }
    return 0;
}

int find_frame(const uint8_t *mp3, int mp3_bytes, int *free_format_bytes, int *ptr_frame_bytes)
{
    for (int pos = 0; pos < mp3_bytes - kHdrSize; pos++)
    {
        const uint8_t *hdr = mp3 + pos;
        if (!hdr_valid(hdr)) continue;
        const int avail = mp3_bytes - pos;
        int frame_bytes = hdr_frame_bytes(hdr, *free_format_bytes);
        int stride = frame_bytes + hdr_padding(hdr);
        if (!frame_bytes)
        {
            const int s = free_format_stride(hdr, avail, &frame_bytes);
            if (s)
            {
                stride = s;
                *free_format_bytes = frame_bytes;
            }
        }
        const bool lined_up = frame_bytes && stride <= avail && followers_agree(hdr, avail, frame_bytes);
        const bool whole_input = pos == 0 && stride == mp3_bytes;
        if (lined_up || whole_input)
        {
            *ptr_frame_bytes = stride;
            return pos;
        }
        *free_format_bytes = 0;
    }
    *ptr_frame_bytes = 0;
    return mp3_bytes;
}

/* ---- bit reader (minimp3.h:248-262) ----------------------------------------------------------- */

uint32_t BitReader::get_slow(int start, int n) const
{
    /* near the end of the frame: gather the (at most 5) bytes that hold the field, then shift it out */
    const int end = start + n;
    const uint8_t *p = buf + (start >> 3);
    const int last = (end - 1) >> 3, first = start >> 3;
    uint64_t acc = 0;
    for (int b = first; b <= last; b++) acc = (acc << 8) | *p++;
    const int tail = 7 - ((end - 1) & 7);
    return (uint32_t)((acc >> tail) & ((n >= 32) ? 0xFFFFFFFFull : ((1ull << n) - 1)));
}

/* ---- side info (minimp3.h:484-607) ------------------------------------------------------------- */

int sample_rate_class(const uint8_t *hdr) { return si_sample_rate_class(hdr); }

int read_side_info(BitReader *bs, GranuleSide *gr, const uint8_t *hdr) { return read_side_info_t<true>(bs, gr, hdr); }
int read_side_info_lite(BitReader *bs, GranuleSide *gr, const uint8_t *hdr) { return read_side_info_lite_t(bs, gr, hdr); }

void fill_gc_desc(GcDesc *d, const GranuleSide &g, const uint8_t *hdr, int ch, int nch, int igr, uint64_t bit_off)
{
    fill_gc_desc_hb(d, g, HdrBits(hdr, nch), ch, igr, bit_off);
}

/* ---- one mp3dec_decode_frame worth of parsing (minimp3.h:1713-1806) ---------------------------- */

void parse_frame(SyncState *st, const uint8_t *mp3, int mp3_bytes, FrameParse *out, bool headers_only, bool lite_side_info)
{
    int i = 0, frame_size = 0;
    out->samples = 0;
    out->reset = out->init_only = false;
    out->layer = 0;
    out->n_gr = out->nch = 0;
    out->success = false;
    out->hdr = nullptr;
    out->payload = nullptr;
    out->payload_bytes = 0;

    if (mp3_bytes > 4 && st->header[0] == 0xff && hdr_compare(st->header, mp3))
    {
        frame_size = hdr_frame_bytes(mp3, st->free_format_bytes) + hdr_padding(mp3);
        if (frame_size != mp3_bytes && (frame_size + kHdrSize > mp3_bytes || !hdr_compare(mp3, mp3 + frame_size)))
            frame_size = 0;
    }
    if (!frame_size)
    {
        *st = SyncState();               /* memset(dec, 0): overlap, QMF history and reservoir are gone */
        out->reset = true;
        i = find_frame(mp3, mp3_bytes, &st->free_format_bytes, &frame_size);
        if (!frame_size || i + frame_size > mp3_bytes)
        {
            out->info.frame_bytes = i;
            return;
        }
    }
    const uint8_t *hdr = mp3 + i;
    memcpy(st->header, hdr, kHdrSize);
    out->hdr = hdr;
    out->info.frame_bytes = i + frame_size;
    out->info.frame_offset = i;
    out->info.channels = hdr_is_mono(hdr) ? 1 : 2;
    out->info.hz = (int)hdr_sample_rate_hz(hdr);
    out->info.layer = hdr_layer(hdr);
    out->info.bitrate_kbps = (int)hdr_bitrate_kbps(hdr);
    out->layer = out->info.layer;
    out->nch = out->info.channels;
    if (headers_only)
    {
        out->samples = (int)hdr_frame_samples(hdr);
        return;
    }

    BitReader bs;
    bs.init(hdr + kHdrSize, frame_size - kHdrSize);
    if (hdr_is_crc(hdr)) bs.get(16);     /* CRC is skipped, never verified */

    if (out->layer != 3)
    {
        /* Layer I/II (minimp3.h:1783-1802): the frame is dropped (and the decoder re-initialised) when its bit
         * allocation asks for more bits than the frame holds */
        out->payload = bs.buf + bs.pos/8;
        out->payload_bytes = (bs.limit - bs.pos)/8;
        BitReader pr;
        pr.init(out->payload, out->payload_bytes);
        L12Info sci;
        l12_read_scale_info(hdr, hdr_bitrate_kbps(hdr), pr, &sci, false);
        const int group = out->layer | 1;
        const long total_bits = (long)pr.pos + 12L*l12_subblock_bits(sci, group);
        out->success = total_bits <= (long)pr.limit;
        if (!out->success)
        {
            st->header[0] = 0;
            out->init_only = true;
            return;
        }
        out->samples = (int)hdr_frame_samples(hdr);
        return;
    }

    out->side_info = bs.buf + bs.pos/8;
    int mdb = lite_side_info ? read_side_info_lite(&bs, out->side, hdr) : read_side_info(&bs, out->side, hdr);
    if (mdb < 0 || bs.pos > bs.limit)
    {
        st->header[0] = 0;               /* mp3dec_init: next call resyncs and clears all state */
        out->init_only = true;
        return;
    }
    out->main_data_begin = mdb;
    out->n_gr = hdr_is_mpeg1(hdr) ? 2 : 1;
    out->payload = bs.buf + bs.pos/8;
    out->payload_bytes = (bs.limit - bs.pos)/8;

    /* L3_restore_reservoir (minimp3.h:1228-1236) + L3_save_reservoir (:1212-1226), byte counts only */
    const int bytes_have = std::min(st->reserv, mdb);
    out->success = st->reserv >= mdb;
    int consumed_bits = 0;
    if (out->success)
        for (int k = 0; k < out->n_gr*out->nch; k++) consumed_bits += out->side[k].part_23_length;
    int remains = bytes_have + out->payload_bytes - (consumed_bits + 7)/8;
    if (remains > kMaxReservoir) remains = kMaxReservoir;
    st->reserv = remains;
    out->samples = out->success ? (int)hdr_frame_samples(hdr) : 0;
}

/* ---- tags (minimp3_ex.h:137-263) ------------------------------------------------------------------ */

void skip_id3v1(const uint8_t *buf, size_t *pbuf_size)
{
    size_t n = *pbuf_size;
    if (n >= 128 && !memcmp(buf + n - 128, "TAG", 3))
    {
        n -= 128;
        if (n >= 227 && !memcmp(buf + n - 227, "TAG+", 4)) n -= 227;
    }
    if (n > 32 && !memcmp(buf + n - 32, "APETAGEX", 8))
    {
        n -= 32;
        const uint8_t *t = buf + n + 8 + 4;
        uint32_t tag_size = ((uint32_t)t[3] << 24) | ((uint32_t)t[2] << 16) | ((uint32_t)t[1] << 8) | t[0];
        if (n >= tag_size) n -= tag_size;
    }
    *pbuf_size = n;
}

size_t skip_id3v2(const uint8_t *buf, size_t buf_size)
{
    if (buf_size >= 10 && !memcmp(buf, "ID3", 3) &&
        !((buf[5] & 15) || (buf[6] & 0x80) || (buf[7] & 0x80) || (buf[8] & 0x80) || (buf[9] & 0x80)))
    {
        size_t n = (((size_t)(buf[6] & 0x7f) << 21) | ((size_t)(buf[7] & 0x7f) << 14) | ((size_t)(buf[8] & 0x7f) << 7) |
                    (size_t)(buf[9] & 0x7f)) + 10;
        if (buf[5] & 16) n += 10;  /* footer */
        return n;
    }
    return 0;
}

void skip_id3(const uint8_t **pbuf, size_t *pbuf_size)
{
    const uint8_t *buf = *pbuf;
    size_t n = *pbuf_size;
    size_t v2 = skip_id3v2(buf, n);
    if (v2)
    {
        if (v2 >= n) v2 = n;
        buf += v2;
        n -= v2;
    }
    skip_id3v1(buf, &n);
    *pbuf = buf;
    *pbuf_size = n;
}

int check_vbrtag(const uint8_t *frame, int frame_size, uint32_t *frames, int *delay, int *padding)
{
    enum { FRAMES_FLAG = 1, BYTES_FLAG = 2, TOC_FLAG = 4, VBR_SCALE_FLAG = 8 };
    if (frame_size < kHdrSize + 8) return 0;
    BitReader bs;
    GranuleSide side[4];
    bs.init(frame + kHdrSize, frame_size - kHdrSize);
    if (hdr_is_crc(frame)) bs.get(16);
    if (read_side_info(&bs, side, frame) < 0) return 0;

    size_t fsz = (size_t)frame_size;
    size_t off = kHdrSize + bs.pos/8;
    if (off > fsz || fsz - off < 8) return 0;
    const uint8_t *tag = frame + off;
    if (memcmp("Xing", tag, 4) && memcmp("Info", tag, 4)) return 0;
    int flags = tag[7];
    if (!(flags & FRAMES_FLAG)) return -1;
    off += 8;
    if (fsz - off < 4) return 0;
    tag = frame + off;
    *frames = ((uint32_t)tag[0] << 24) | ((uint32_t)tag[1] << 16) | ((uint32_t)tag[2] << 8) | tag[3];
    off += 4;
    if (flags & BYTES_FLAG) { if (fsz - off < 4) return 0; off += 4; }
    if (flags & TOC_FLAG) { if (fsz - off < 100) return 0; off += 100; }
    if (flags & VBR_SCALE_FLAG) { if (fsz - off < 4) return 0; off += 4; }
    *delay = *padding = 0;
    if (fsz - off < 1) return 0;
    tag = frame + off;
    if (*tag)
    {   /* encoder extension (LAME, Lavc, ...) */
        if (fsz - off <= 35) return 0;
        off += 21;
        tag = frame + off;
        *delay = ((tag[0] << 4) | (tag[1] >> 4)) + (528 + 1);
        *padding = (((tag[1] & 0xF) << 8) | tag[2]) - (528 + 1);
    }
    return 1;
}

int detect_buf(const uint8_t *buf, size_t buf_size)
{
    if (!buf || (size_t)-1 == buf_size) return -1;      /* MP3D_E_PARAM */
    size_t filled = buf_size;
    if (filled < 10) return -4;                          /* MP3D_E_USER: too small */
    if (skip_id3v2(buf, filled)) return 0;
    skip_id3v1(buf, &filled);
    if (filled > 16*1024) filled = 16*1024;              /* MINIMP3_BUF_SIZE */
    int free_format_bytes = 0, frame_size = 0;
    find_frame(buf, (int)filled, &free_format_bytes, &frame_size);
    return frame_size ? 0 : -4;
}

/* ---- the head of a stream as mp3dec_load_buf sees it (minimp3_ex.h:313-371) -------------------------------------
 * Tags are stepped over, the first frame fixes the stream parameters, and a Layer III first frame that is a
 * Xing / Info tag is dropped -- with a frame count it also tells the length: frames x samples, minus the encoder
 * delay in front (to_skip) and the padding behind. */
bool open_stream_head(const uint8_t **pbuf, size_t *psize, StreamHead *head)
{
    const uint8_t *buf = *pbuf;
    size_t size = *psize;
    skip_id3(&buf, &size);
    int frame_size = 0;
    while (size && !frame_size)
    {
        int free_format_bytes = 0;
        const int junk = find_frame(buf, (int)std::min(size, (size_t)INT_MAX), &free_format_bytes, &frame_size);
        buf += junk;
        size -= (size_t)junk;
        if (!junk && !frame_size) break;       /* nothing left to search */
    }
    *pbuf = buf;
    *psize = size;
    if (!frame_size) return false;
    FrameInfo &f = head->first;
    f.channels = hdr_is_mono(buf) ? 1 : 2;
    f.hz = (int)hdr_sample_rate_hz(buf);
    f.layer = hdr_layer(buf);
    f.bitrate_kbps = (int)hdr_bitrate_kbps(buf);
    f.frame_bytes = frame_size;
    f.frame_offset = 0;
    head->frame_samples = (int)hdr_frame_samples(buf);
    head->to_skip = head->detected_samples = 0;
    if (f.layer != 3) return true;
    uint32_t n_frames = 0;
    int delay = 0, padding = 0;
    const int tag = check_vbrtag(buf, frame_size, &n_frames, &delay, &padding);
    if (tag > 0)
    {
        const uint64_t per_frame = (uint64_t)hdr_frame_samples(buf)*(uint64_t)f.channels;
        uint64_t total = per_frame*n_frames;
        head->to_skip = (uint64_t)((int64_t)delay*f.channels);
        if (total >= head->to_skip) total -= head->to_skip;
        const int64_t pad = (int64_t)padding*f.channels;
        if (pad > 0 && total >= (uint64_t)pad) total -= (uint64_t)pad;
        if (!total) return false;              /* a tag that announces nothing */
        head->detected_samples = total;
    }
    if (tag)
    {
        *pbuf = buf + frame_size;
        *psize = size - (size_t)frame_size;
    }
    return true;
}

/* ---- whole-stream walk (minimp3_ex.h:308-525 in buffer mode) ----------------------------------------- */

static void build_tiles(StreamPlan *plan, int tile_granules)
{
    int32_t chain[2][2] = { { -1, -1 }, { -1, -1 } };   /* [ch][0] = previous gc, [ch][1] = the one before */
    uint64_t pcm_off = 0;
    size_t n = plan->granules.size();
    size_t g = 0;
    int last_kind = -1;
    while (g < n)
    {
        const GranuleRef &first = plan->granules[g];
        /* a layer change always comes with a decoder reset (hdr_compare checks the layer bits) */
        if (first.reset || first.kind != last_kind)
            chain[0][0] = chain[0][1] = chain[1][0] = chain[1][1] = -1;
        last_kind = first.kind;
        const uint32_t per_gr = first.kind ? 384u : 576u;
        TileDesc t;
        memset(&t, 0, sizeof(t));
        t.gc_first = first.gc;
        t.nch = first.nch;
        t.pcm_off = pcm_off;
        for (int c = 0; c < 2; c++) { t.halo[c][0] = chain[c][0]; t.halo[c][1] = chain[c][1]; }
        int cnt = 0;
        /* Layer III mono tiles take twice the granules (MP3B_TILE_GRANULES_MONO); Layer I/II tiles keep the base size */
        const int cap = (!first.kind && first.nch == 1) ? tile_granules*(MP3B_TILE_GRANULES_MONO/MP3B_TILE_GRANULES) : tile_granules;
        while (g < n && cnt < cap)
        {
            const GranuleRef &r = plan->granules[g];
            if (cnt && (r.reset || r.kind != first.kind || r.nch != first.nch || r.gc != first.gc + (uint32_t)(cnt*first.nch))) break;
            for (int c = 0; c < r.nch; c++)
            {
                chain[c][1] = chain[c][0];
                chain[c][0] = (int32_t)(r.gc + c);
            }
            pcm_off += per_gr*r.nch;
            cnt++;
            g++;
        }
        t.n_gr = (uint8_t)cnt;
        if (first.kind) plan->tiles12.push_back(t);
        else plan->tiles.push_back(t);
    }
}

/* Lane order of the entropy kernel: granule-channels of similar main-data size next to each other, so the
 * 32 lanes of a warp run about the same number of Huffman iterations (counting sort on part_23_length). */
static void build_order(StreamPlan *plan)
{
    uint32_t count[257];
    memset(count, 0, sizeof(count));
    if (!plan->gc_key.empty() || plan->gcs.empty())
    {
        /* device_side_info: sizes were kept per descriptor slot, ~0 marks the padding slots */
        const size_t n = plan->gc_key.size();
        for (size_t i = 0; i < n; i++)
            if (plan->gc_key[i] != 0xFFFFFFFFu) count[((plan->gc_key[i] & 0xFFFF) >> 4) + 1]++;
        for (int b = 0; b < 256; b++) count[b + 1] += count[b];
        plan->order.resize(count[256]);
        for (size_t i = 0; i < n; i++)
            if (plan->gc_key[i] != 0xFFFFFFFFu) plan->order[count[(plan->gc_key[i] & 0xFFFF) >> 4]++] = (uint32_t)i;
        return;
    }
    const size_t n = plan->gcs.size();
    std::vector<uint8_t> real(n, 0);
    for (const GranuleRef &g : plan->granules)
        if (!g.kind)
            for (int c = 0; c < g.nch; c++) real[g.gc + c] = 1;
    for (size_t i = 0; i < n; i++)
        if (real[i]) count[(plan->gcs[i].part_23_length >> 4) + 1]++;
    for (int b = 0; b < 256; b++) count[b + 1] += count[b];
    plan->order.resize(count[256]);
    for (size_t i = 0; i < n; i++)
        if (real[i]) plan->order[count[plan->gcs[i].part_23_length >> 4]++] = (uint32_t)i;
}

void walk_stream(const uint8_t *buf, size_t size, uint8_t *md_dst, uint64_t md_base_bits,
                 const WalkOptions &opt, StreamPlan *plan)
{
    *plan = StreamPlan();
    uint64_t detected_samples = 0;
    uint64_t to_skip = 0;
    FrameInfo first{};
    if (!opt.raw_frames)
    {
        StreamHead head;
        if (!open_stream_head(&buf, &size, &head)) return;
        first = head.first;
        to_skip = head.to_skip;
        detected_samples = head.detected_samples;
        plan->channels = first.channels;
        plan->hz = first.hz;
        plan->layer = first.layer;
    }

    {
        /* one allocation per array for the common constant-bit-rate case instead of log2(n) regrowths */
        const size_t guess = size/(size_t)std::max(first.frame_bytes, 96) + 8;
        if (opt.device_side_info) { plan->frame_recs.reserve(guess); plan->gc_key.reserve(guess*4); }
        else plan->gcs.reserve(guess*4);
        plan->granules.reserve(guess*3);
    }
    SyncState st;
    FrameParse fp;
    memset(&fp.info, 0, sizeof(fp.info));
    uint64_t md_len = 0;                 /* bytes of main data written so far (|C|) */
    uint64_t delivered = 0;              /* mp3dec_file_info_t.samples */
    uint64_t skip_left = to_skip;
    uint64_t bitrate_sum = 0, frames_counted = 0;
    bool pending_reset = true;           /* first decoded granule starts from zero state */
    do
    {
        parse_frame(&st, buf, (int)std::min(size, (size_t)INT_MAX), &fp, false, opt.device_side_info);
        plan->frames++;
        if (fp.reset || fp.init_only) pending_reset = true;
        if (fp.reset) { /* reservoir emptied: earlier main data can no longer be referenced */ }
        if (fp.layer == 3 && fp.payload && !fp.init_only && fp.n_gr)
        {
            /* frame's main data begins main_data_begin bytes before its own payload in C */
            if (fp.success)
            {
                uint64_t start_bits = (md_len - (uint64_t)fp.main_data_begin)*8u;
                bool stop = false;
                if (!opt.raw_frames)
                {
                    if (plan->hz != fp.info.hz || plan->layer != fp.info.layer) { plan->ret = -5; stop = true; }
                    else if (plan->channels && plan->channels != fp.info.channels)
                    {
                        if (opt.allow_mono_stereo) plan->channels = 0;
                        else { plan->ret = -5; stop = true; }
                    }
                }
                if (stop) break;
                uint64_t bit = md_base_bits + start_bits;
                const bool istereo = (fp.hdr[3] & 0x10) && fp.nch == 2;
                const HdrBits hb(fp.hdr, fp.nch);
                FrameRec rec;
                bool padded = false;
                for (int igr = 0; igr < fp.n_gr; igr++)
                {
                    if (fp.nch == 2 && (plan->n_gc & 1))
                    {
                        /* keep stereo pairs on even indices (MS / intensity partners are gc ^ 1) */
                        padded = true;
                        if (opt.device_side_info) plan->gc_key.push_back(0xFFFFFFFFu);
                        else
                        {
                            GcDesc pad;
                            memset(&pad, 0, sizeof(pad));
                            pad.flags1 = GCF1_MPEG1;
                            plan->gcs.push_back(pad);
                        }
                        plan->n_gc++;
                    }
                    GranuleRef gref;
                    gref.gc = plan->n_gc;
                    gref.nch = (uint8_t)fp.nch;
                    gref.reset = pending_reset ? 1 : 0;
                    gref.kind = 0;
                    pending_reset = false;
                    plan->granules.push_back(gref);
                    if (istereo) plan->istereo_gcs.push_back(gref.gc);
                    if (igr == 0) rec.gc_first = gref.gc;
                    for (int ch = 0; ch < fp.nch; ch++)
                    {
                        const GranuleSide &g = fp.side[igr*fp.nch + ch];
                        if (opt.device_side_info) plan->gc_key.push_back((uint32_t)g.big_values << 16 | g.part_23_length);
                        else
                        {
                            GcDesc d;
                            fill_gc_desc_hb(&d, g, hb, ch, igr, bit);
                            plan->gcs.push_back(d);
                        }
                        bit += g.part_23_length;
                        plan->n_gc++;
                    }
                }
                if (opt.device_side_info)
                {
                    /* the GPU parses the side info itself (k_l3_side_info): hand it the bytes */
                    rec.md_bit = md_base_bits + start_bits;
                    memcpy(rec.hdr, fp.hdr, 4);
                    rec.nch = (uint8_t)fp.nch;
                    rec.n_gr = (uint8_t)fp.n_gr;
                    memset(rec.pad, 0, sizeof(rec.pad));
                    rec.pad[0] = padded ? 1 : 0;       /* one slot was skipped right in front of granule 0 */
                    const size_t room = (size_t)(fp.payload + fp.payload_bytes - fp.side_info);
                    memcpy(rec.si, fp.side_info, std::min<size_t>(32, room));
                    if (room < 32) memset(rec.si + room, 0, 32 - room);
                    memset(rec.pad2, 0, sizeof(rec.pad2));
                    plan->frame_recs.push_back(rec);
                }
                plan->frames_decoded++;
                if (opt.raw_frames && !plan->hz)
                {
                    plan->hz = fp.info.hz; plan->layer = 3; plan->channels = fp.info.channels;
                }
                uint64_t s = (uint64_t)fp.samples*(uint64_t)fp.info.channels;
                plan->samples_decoded += s;
                uint64_t skip = std::min(s, skip_left);
                skip_left -= skip;
                delivered += s - skip;
                bitrate_sum += (uint64_t)fp.info.bitrate_kbps;
                frames_counted++;
            }
            /* payload joins C whether or not the frame was decodable (L3_save_reservoir runs either way) */
            memcpy(md_dst + md_len, fp.payload, (size_t)fp.payload_bytes);
            md_len += (uint64_t)fp.payload_bytes;
        } else if (fp.layer && fp.layer != 3 && fp.samples)
        {
            bool stop = false;
            if (!opt.raw_frames)
            {
                if (plan->hz != fp.info.hz || plan->layer != fp.info.layer) { plan->ret = -5; stop = true; }
                else if (plan->channels && plan->channels != fp.info.channels)
                {
                    if (opt.allow_mono_stereo) plan->channels = 0;
                    else { plan->ret = -5; stop = true; }
                }
            }
            if (stop) break;
            L12FrameDesc fd;
            memset(&fd, 0, sizeof(fd));
            fd.bit_off = md_base_bits + md_len*8u;
            fd.block_first = plan->n_blockch;
            memcpy(fd.hdr, fp.hdr, 4);
            fd.kbps = (uint16_t)fp.info.bitrate_kbps;
            fd.nch = (uint8_t)fp.nch;
            fd.n_blocks = (uint8_t)(fp.layer == 1 ? 1 : 3);
            plan->l12_frames.push_back(fd);
            for (int b = 0; b < fd.n_blocks; b++)
            {
                GranuleRef gref;
                gref.gc = plan->n_blockch;
                gref.nch = (uint8_t)fp.nch;
                gref.reset = pending_reset ? 1 : 0;
                gref.kind = 1;
                pending_reset = false;
                plan->granules.push_back(gref);
                plan->n_blockch += (uint32_t)fp.nch;
            }
            memcpy(md_dst + md_len, fp.payload, (size_t)fp.payload_bytes);
            md_len += ((uint64_t)fp.payload_bytes + 3u) & ~3ull;     /* keep every frame payload word-aligned */
            plan->frames_decoded++;
            if (opt.raw_frames && !plan->hz)
            {
                plan->hz = fp.info.hz; plan->layer = fp.info.layer; plan->channels = fp.info.channels;
            }
            uint64_t s = (uint64_t)fp.samples*(uint64_t)fp.info.channels;
            plan->samples_decoded += s;
            uint64_t skip = std::min(s, skip_left);
            skip_left -= skip;
            delivered += s - skip;
            bitrate_sum += (uint64_t)fp.info.bitrate_kbps;
            frames_counted++;
        }
        buf += fp.info.frame_bytes;
        size -= (size_t)fp.info.frame_bytes;
    } while (fp.info.frame_bytes);

    plan->main_data_bytes = md_len;
    plan->out_skip = to_skip - skip_left;
    if (detected_samples && delivered > detected_samples) delivered = detected_samples;
    plan->out_samples = delivered;
    if (frames_counted) plan->avg_bitrate_kbps = (int)(bitrate_sum/frames_counted);
    build_tiles(plan, opt.tile_granules);
    build_order(plan);
}

}  // namespace mp3b
