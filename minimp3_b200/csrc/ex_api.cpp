/*
 * ex_api.cpp -- the iterate / streaming / seek entry points of minimp3_ex.h on top of this library's batch and frame
 * APIs (every PCM sample comes out of the CUDA kernels: look-ahead windows through mp3dec_decode_batch_cuda, single
 * frames through mp3dec_decode_frame).
 *
 * Behaviour follows /root/reference/minimp3_ex.h -- mp3dec_iterate_buf (:527-564), mp3dec_load_index (:641-698),
 * mp3dec_ex_open_buf (:700-717), mp3dec_ex_seek incl. the bit-reservoir pre-roll (:743-893), mp3dec_ex_read_frame /
 * mp3dec_ex_read (:895-1013), the stdio flavours (:1199-1434) -- written here around this library's own pieces: one
 * frame walk (walk_frames) under both the iterate call and the index builder, an index pass that asks the parser's
 * decoder-state emulation (not the decoder) which frames produce output, look-ahead windows under the reads.
 * The *_cb flavours read the whole stream through the caller's callbacks into memory first and then run the
 * buffer flavours (same results; I/O happens up front instead of lazily).
 */
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "host_parse.h"

#include "../../include/minimp3_b200.h"

using namespace mp3b;

extern "C" int mp3dec_decode_frame_f32(mp3dec_t *dec, const uint8_t *mp3, int mp3_bytes, float *pcm, mp3dec_frame_info_t *info);
extern "C" __attribute__((visibility("hidden"))) int mp3dec_decode_frame_direct(mp3dec_t *dec, const uint8_t *mp3, int mp3_bytes, void *pcm,
                                                                                mp3dec_frame_info_t *info, int float_output);

namespace {

template <class S>
struct ExT   /* mp3dec_ex_t with sample type S (minimp3_ex.h:74-88) */
{
    mp3dec_t mp3d;
    mp3dec_map_info_t file;
    mp3dec_io_t *io;
    mp3dec_index_t index;
    uint64_t offset, samples, detected_samples, cur_sample, start_offset, end_offset;
    mp3dec_frame_info_t info;
    S buffer[MINIMP3_MAX_SAMPLES_PER_FRAME];
    size_t input_consumed, input_filled;
    int is_file, flags, vbr_tag_found, indexes_built;
    int free_format_bytes;
    int buffer_samples, buffer_consumed, to_skip, start_delay;
    int last_error;
};
static_assert(sizeof(ExT<int16_t>) == sizeof(mp3dec_ex_t), "mp3dec_ex_t layout");

/* the streaming API keeps its own look-ahead windows: its single-frame calls go straight to the GPU frame path */
inline int decode_frame(mp3dec_t *d, const uint8_t *p, int n, int16_t *pcm, mp3dec_frame_info_t *info)
{
    return mp3dec_decode_frame_direct(d, p, n, pcm, info, 0);
}
inline int decode_frame(mp3dec_t *d, const uint8_t *p, int n, float *pcm, mp3dec_frame_info_t *info)
{
    return mp3dec_decode_frame_direct(d, p, n, pcm, info, 1);
}

void fill_info(mp3dec_frame_info_t *fi, const uint8_t *hdr, int frame_size)
{
    fi->channels = hdr_is_mono(hdr) ? 1 : 2;
    fi->hz = (int)hdr_sample_rate_hz(hdr);
    fi->layer = hdr_layer(hdr);
    fi->bitrate_kbps = (int)hdr_bitrate_kbps(hdr);
    fi->frame_bytes = frame_size;
}

/* ---- frame walk shared by mp3dec_iterate_buf and the index builder ------------------------------------------
 * Tags are skipped, then every frame the sync search finds (minimp3_ex.h:527-564 defines which) is handed to
 * visit(frame, frame_size, free_format_bytes, bytes_left, offset, info); a non-zero return ends the walk and is
 * returned. */
template <class Visit>
int walk_frames(const uint8_t *buf, size_t size, Visit &&visit)
{
    const uint8_t *const origin = buf;
    skip_id3(&buf, &size);
    mp3dec_frame_info_t fi;
    memset(&fi, 0, sizeof(fi));
    while (size)
    {
        int free_format_bytes = 0, frame_size = 0;
        const int junk = find_frame(buf, (int)std::min(size, (size_t)INT_MAX), &free_format_bytes, &frame_size);
        buf += junk;
        size -= (size_t)junk;
        if (!frame_size)
        {
            if (junk) continue;      /* the search window (INT_MAX bytes) ended without a frame: go on behind it */
            break;
        }
        fill_info(&fi, buf, frame_size);
        const int r = visit(buf, frame_size, free_format_bytes, size, (uint64_t)(buf - origin), &fi);
        if (r) return r;
        buf += frame_size;
        size -= (size_t)frame_size;
    }
    return 0;
}

/* Length and encoder delay from a Xing/Info tag (minimp3_ex.h:655-676): samples the stream delivers after the
 * decoder delay at the front and the padding at the end are cut. */
struct TagLength
{
    uint64_t samples;
    int delay;          /* channel-inclusive samples to drop at the start */
};

TagLength length_from_tag(const uint8_t *frame, int channels, uint32_t n_frames, int delay, int padding)
{
    TagLength t;
    t.delay = delay*channels;
    t.samples = (uint64_t)hdr_frame_samples(frame)*(uint64_t)channels*(uint64_t)n_frames;
    if (t.samples >= (uint64_t)t.delay) t.samples -= (uint64_t)t.delay;
    const int64_t pad = (int64_t)padding*channels;
    if (pad > 0 && t.samples >= (uint64_t)pad) t.samples -= (uint64_t)pad;
    return t;
}

/* ---- stream index (mp3dec_load_index, minimp3_ex.h:641-698) -------------------------------------------------
 * One pass over the frames of buf: the first frame fixes the stream parameters and may carry a VBR tag (which ends
 * the pass: the length is known); every other frame becomes an index entry (offset, samples delivered before it).
 * How many samples a frame delivers: the reference decodes frames until the first one that produces output (a cut
 * file starts with frames whose bit reservoir is missing) and counts whole frames from then on.  Whether a frame
 * produces output is a property of the bitstream alone, so the parser's decoder-state emulation (parse_frame)
 * answers it here without touching the GPU. */
template <class S>
class IndexPass
{
public:
    explicit IndexPass(ExT<S> *d) : dec_(d) {}

    int frame(const uint8_t *hdr, int frame_size, int free_format_bytes, size_t bytes_left, uint64_t offset,
              const mp3dec_frame_info_t *info)
    {
        ExT<S> &d = *dec_;
        if (!have_params_ && !d.index.frames && !d.start_offset)
        {
            have_params_ = true;
            d.info = *info;
            d.start_offset = d.offset = offset;
            d.end_offset = offset + bytes_left;
            d.free_format_bytes = free_format_bytes;
            if (info->layer == 3)
            {
                uint32_t n_frames = 0;
                int delay = 0, padding = 0;
                const int tag = check_vbrtag(hdr, frame_size, &n_frames, &delay, &padding);
                if (tag) d.start_offset = d.offset = offset + (uint64_t)frame_size;     /* the tag frame holds no audio */
                if (tag > 0)
                {
                    const TagLength t = length_from_tag(hdr, info->channels, n_frames, delay, padding);
                    d.start_delay = d.to_skip = t.delay;
                    d.samples = d.detected_samples = t.samples;
                    d.vbr_tag_found = 1;
                    return MP3D_E_USER;
                }
                if (tag < 0) return 0;       /* tag without a frame count: skipped, the scan goes on */
            }
        }
        if (d.flags & MP3D_DO_NOT_SCAN) return MP3D_E_USER;
        mp3dec_frame_t e;
        e.offset = offset;
        e.sample = d.samples;
        entries_.push_back(e);
        uint64_t produced = (uint64_t)hdr_frame_samples(hdr);
        if (!counting_whole_frames_ && entries_.size() < 256)
        {
            FrameParse fp;
            memset(&fp.info, 0, sizeof(fp.info));
            parse_frame(&probe_, hdr, (int)std::min(bytes_left, (size_t)INT_MAX), &fp);
            produced = (uint64_t)fp.samples;
            counting_whole_frames_ = fp.samples != 0;
        } else
            counting_whole_frames_ = true;
        d.samples += produced*(uint64_t)info->channels;
        return 0;
    }

    /* hands the entries to the public index (malloc()ed: mp3dec_ex_close frees it) */
    int finish(uint64_t offset_bias)
    {
        ExT<S> &d = *dec_;
        if (entries_.empty()) return 0;
        mp3dec_frame_t *out = (mp3dec_frame_t *)malloc(entries_.size()*sizeof(mp3dec_frame_t));
        if (!out) return MP3D_E_MEMORY;
        for (size_t k = 0; k < entries_.size(); k++)
        {
            out[k] = entries_[k];
            out[k].offset += offset_bias;
        }
        free(d.index.frames);
        d.index.frames = out;
        d.index.num_frames = d.index.capacity = entries_.size();
        return 0;
    }

private:
    ExT<S> *dec_;
    std::vector<mp3dec_frame_t> entries_;
    SyncState probe_;                      /* decoder-state emulation for the "does it produce output" question */
    bool have_params_ = false;
    bool counting_whole_frames_ = false;
};

template <class S>
int scan_stream(ExT<S> *dec, const uint8_t *buf, size_t size, uint64_t offset_bias)
{
    IndexPass<S> pass(dec);
    const int r = walk_frames(buf, size, [&](const uint8_t *f, int fs, int ffb, size_t left, uint64_t off, mp3dec_frame_info_t *fi) {
        return pass.frame(f, fs, ffb, left, off, fi);
    });
    if (r && r != MP3D_E_USER) return r;
    return pass.finish(offset_bias);
}

/* index entry decoding has to start from for `position`: the entry holding it -- the last one that starts before
 * it, or the one starting exactly there (minimp3_ex.h:720-740) */
size_t entry_for_sample(const mp3dec_index_t &idx, uint64_t position)
{
    const mp3dec_frame_t *first = idx.frames, *last = idx.frames + idx.num_frames;
    const mp3dec_frame_t *it = std::lower_bound(first, last, position,
                                                [](const mp3dec_frame_t &f, uint64_t pos) { return f.sample < pos; });
    if (it != last && it->sample == position) return (size_t)(it - first);
    return it == first ? 0 : (size_t)(it - first) - 1;
}

void drop_window(const void *owner);

template <class S>
int ex_open_buf(ExT<S> *dec, const uint8_t *buf, size_t buf_size, int flags)
{
    if (!dec || !buf || (size_t)-1 == buf_size || (flags & (~MP3D_FLAGS_MASK))) return MP3D_E_PARAM;
    drop_window(dec);            /* the object may be reused without a close */
    memset(dec, 0, sizeof(*dec));
    dec->file.buffer = buf;
    dec->file.size = buf_size;
    dec->flags = flags;
    const int ret = scan_stream(dec, buf, buf_size, 0);
    if (ret) return ret;
    mp3dec_init(&dec->mp3d);
    dec->buffer_samples = 0;
    dec->indexes_built = !(dec->vbr_tag_found || (flags & MP3D_DO_NOT_SCAN));
    dec->flags &= ~MP3D_DO_NOT_SCAN;
    return 0;
}

/* payload bytes (what follows header, CRC and side info) of the Layer III frame at hdr, or -1 if its side info is
 * not valid: what the frame adds to the bit reservoir */
int l3_payload_bytes(const uint8_t *hdr, int free_format_bytes)
{
    const int frame_size = hdr_frame_bytes(hdr, free_format_bytes) + hdr_padding(hdr);
    BitReader bs;
    GranuleSide side[4];
    bs.init(hdr + 4, frame_size - 4);
    if (hdr_is_crc(hdr)) bs.get(16);
    if (read_side_info(&bs, side, hdr) < 0) return -1;
    return (bs.limit - bs.pos)/8;
}

/* everything a seek leaves behind except the position: empty frame buffer, no sticky error, fresh decoder */
template <class S>
void restart_at(ExT<S> *dec, uint64_t offset, int to_skip)
{
    dec->offset = offset;
    dec->to_skip = to_skip;
    dec->buffer_samples = dec->buffer_consumed = 0;
    dec->input_consumed = dec->input_filled = 0;
    dec->last_error = 0;
    mp3dec_init(&dec->mp3d);
}

/* mp3dec_ex_seek (minimp3_ex.h:743-893), buffer mode.  Byte positions are taken as they are.  Sample positions go
 * through the index: decoding restarts MINIMP3_PREDECODE_FRAMES in front of the frame that holds the sample -- for
 * Layer III further back, until the frames in between can refill the 511-byte bit reservoir -- and the samples
 * decoded on the way are dropped (to_skip). */
template <class S>
int ex_seek(ExT<S> *dec, uint64_t position)
{
    if (!dec) return MP3D_E_PARAM;
    if (!(dec->flags & MP3D_SEEK_TO_SAMPLE))
    {
        dec->cur_sample = 0;
        restart_at(dec, std::min<uint64_t>(position, dec->file.size), dec->to_skip);
        return 0;
    }
    if (dec->samples && position > dec->samples) position = dec->samples;
    dec->cur_sample = position;
    const uint64_t target = position > UINT64_MAX - (uint64_t)dec->start_delay ? UINT64_MAX : position + (uint64_t)dec->start_delay;
    if (target && !dec->indexes_built)
    {
        /* opened through a VBR tag or with MP3D_DO_NOT_SCAN: the index is built now, from the first audio frame */
        dec->indexes_built = 1;
        dec->samples = 0;
        dec->buffer_samples = 0;
        const int ret = scan_stream(dec, dec->file.buffer + dec->start_offset, (size_t)(dec->file.size - dec->start_offset), dec->start_offset);
        if (ret) return ret;
        dec->samples = dec->detected_samples;
    }
    if (!target || !dec->index.frames)
    {
        restart_at(dec, dec->start_offset, 0);
        return 0;
    }
    const mp3dec_frame_t *entries = dec->index.frames;
    size_t at = entry_for_sample(dec->index, target);
    if (at)
    {
        at -= std::min(at, (size_t)MINIMP3_PREDECODE_FRAMES);
        if (dec->info.layer == 3)
        {
            int reservoir_gap = 511;
            while (at && reservoir_gap > 0)
            {
                at--;
                const int payload = l3_payload_bytes(dec->file.buffer + entries[at].offset, dec->free_format_bytes);
                if (payload < 0) break;
                reservoir_gap -= payload;
            }
        }
    }
    const uint64_t restart_offset = entries[at].offset;
    uint64_t drop = target > entries[at].sample ? target - entries[at].sample : 0;
    /* leading entries that delivered nothing (their successor still starts at sample 0) are decoded again on the way
     * and will deliver whole frames then */
    for (; at + 1 < dec->index.num_frames && !entries[at].sample && !entries[at + 1].sample; at++)
        drop += (uint64_t)hdr_frame_samples(dec->file.buffer + entries[at].offset)*(uint64_t)dec->info.channels;
    restart_at(dec, restart_offset, (int)std::min<uint64_t>(drop, (uint64_t)INT_MAX));
    return 0;
}


/* ---- look-ahead windows for the streaming API ---------------------------------------------------------------
 * mp3dec_ex_read* decodes frame after frame from dec->offset, always starting from a freshly initialised decoder
 * (open and every seek end in mp3dec_init, minimp3_ex.h:717, 892).  Instead of one GPU round trip per frame, the
 * frames of a window are decoded in ONE batch call in raw-frame mode -- which is defined as exactly that frame
 * loop -- and handed out one by one.  When a window is used up the next one starts a few frames early (preroll_frames) with a
 * fresh decoder and drops their output: the overlap / synthesis state of the decoder only reaches one granule
 * back and the bit reservoir at most 511 bytes, so from there on the samples are those of the uninterrupted loop
 * (the same argument the reference's own seek relies on, minimp3_ex.h:826-871). */
struct WindowFrame
{
    uint64_t off;               /* file offset the frame call starts at */
    mp3dec_frame_info_t info;   /* what mp3dec_decode_frame reports for it */
    int samples;                /* per channel, 0 if nothing was produced */
    bool found;                 /* false: no frame in the rest of the input, only info.frame_bytes is reported */
    size_t pcm_pos;             /* first sample in the window's PCM */
};

struct Window
{
    std::vector<WindowFrame> frames;
    void *pcm = nullptr;        /* malloc()ed by the batch call */
    size_t next = 0;            /* next frame to hand out */
    ~Window() { free(pcm); }
};

/* Frames a window that starts mid-stream decodes again (and drops) so that the decoder state at its first
 * delivered frame is that of the uninterrupted loop: the overlap / synthesis history reaches one granule back, the
 * Layer III bit reservoir 511 BYTES -- at 8 kbit/s that is dozens of frames -- so the count follows the payload
 * bytes (what mp3dec_ex_seek does, minimp3_ex.h:826-871), with kPrerollMin frames at least. */
constexpr size_t kPrerollMin = 4, kPrerollMax = 256;

/* frames per window; MP3B_DEBUG_WINDOW_FRAMES shrinks it so that tests cross many window boundaries */
size_t window_frames()
{
    const char *e = getenv("MP3B_DEBUG_WINDOW_FRAMES");
    const long v = e ? atol(e) : 0;
    return v > 0 ? (size_t)v : (size_t)4096;
}

/* how many of frames[0 .. end) to decode again in front of frames[end]; base = the bytes at file offset base_off */
size_t preroll_frames(const std::vector<WindowFrame> &frames, size_t end, const uint8_t *base, uint64_t base_off)
{
    size_t n = 0;
    int reservoir_gap = 511;
    while (n < end && n < kPrerollMax && (n < kPrerollMin || reservoir_gap > 0))
    {
        const WindowFrame &f = frames[end - 1 - n];
        n++;
        if (!f.found || f.info.layer != 3) continue;      /* nothing found / no reservoir in Layers I and II */
        const uint8_t *hdr = base + (f.off - base_off) + f.info.frame_offset;
        const int side = hdr_is_mpeg1(hdr) ? (f.info.channels == 1 ? 17 : 32) : (f.info.channels == 1 ? 9 : 17);
        reservoir_gap -= std::max(0, f.info.frame_bytes - f.info.frame_offset - 4 - (hdr_is_crc(hdr) ? 2 : 0) - side);
    }
    return n;
}

std::mutex g_win_mu;
std::unordered_map<const void *, Window *> g_windows;

void drop_window(const void *owner)
{
    std::lock_guard<std::mutex> lock(g_win_mu);
    auto it = g_windows.find(owner);
    if (it != g_windows.end()) { delete it->second; g_windows.erase(it); }
}

Window *find_window(const void *owner)
{
    std::lock_guard<std::mutex> lock(g_win_mu);
    auto it = g_windows.find(owner);
    return it == g_windows.end() ? nullptr : it->second;
}

/* decodes the frames of buf[start, end) that a fresh decoder would, at most max_frames of them */
template <class S>
Window *build_window(const uint8_t *buf, uint64_t start, uint64_t end, size_t max_frames)
{
    Window *w = new Window();
    SyncState st;
    FrameParse fp;
    memset(&fp.info, 0, sizeof(fp.info));
    uint64_t off = start;
    size_t pcm_pos = 0;
    while (off < end && w->frames.size() < max_frames)
    {
        parse_frame(&st, buf + off, (int)std::min<uint64_t>(end - off, (uint64_t)INT_MAX), &fp);
        if (!fp.info.frame_bytes) break;
        WindowFrame f;
        f.off = off;
        f.info.frame_bytes = fp.info.frame_bytes; f.info.frame_offset = fp.info.frame_offset;
        f.info.channels = fp.info.channels; f.info.hz = fp.info.hz; f.info.layer = fp.info.layer;
        f.info.bitrate_kbps = fp.info.bitrate_kbps;
        f.samples = fp.samples;
        f.found = fp.hdr != nullptr;
        f.pcm_pos = pcm_pos;
        pcm_pos += (size_t)fp.samples*(size_t)fp.info.channels;
        w->frames.push_back(f);
        off += (uint64_t)fp.info.frame_bytes;
    }
    if (w->frames.empty()) { delete w; return nullptr; }
    mp3dec_batch_stream_t s = { buf + start, (size_t)(off - start) };
    mp3dec_batch_opts_t o;
    memset(&o, 0, sizeof(o));
    o.float_output = sizeof(S) == 4;
    o.raw_frames = 1;
    o.output = MP3D_BATCH_OUT_MALLOC;
    o.host_threads = 1;
    mp3dec_file_info_t fi;
    memset(&fi, 0, sizeof(fi));
    int ret = 0;
    if (mp3dec_decode_batch_cuda(&s, 1, &fi, &ret, &o) || ret || fi.samples != pcm_pos)
    {
        free(fi.buffer);
        delete w;
        return nullptr;       /* caller falls back to the frame-by-frame path */
    }
    w->pcm = fi.buffer;
    return w;
}

/* mp3dec_decode_frame(&dec->mp3d, file + dec->offset, ...) served from a look-ahead window */
template <class S>
int window_decode_frame(ExT<S> *dec, uint64_t end_offset, S *pcm, mp3dec_frame_info_t *info)
{
    const uint8_t *buf = dec->file.buffer;
    Window *w = find_window(dec);
    const bool fresh = dec->mp3d.header[0] == 0;      /* mp3dec_init since the last frame: open or seek */
    if (w && !fresh && w->next < w->frames.size() && w->frames[w->next].off == dec->offset)
    {
        /* the usual case: the next frame of the window */
    } else if (w && !fresh && w->next == w->frames.size() && w->next > 0 &&
               w->frames.back().off + (uint64_t)w->frames.back().info.frame_bytes == dec->offset)
    {
        /* window used up: the next one re-decodes the frames that rebuild the decoder state */
        const size_t back = preroll_frames(w->frames, w->frames.size(), buf, 0);
        const uint64_t restart = w->frames[w->frames.size() - back].off;
        Window *nw = build_window<S>(buf, restart, end_offset, window_frames() + back);
        drop_window(dec);
        w = nullptr;
        if (nw)
        {
            size_t k = 0;
            while (k < nw->frames.size() && nw->frames[k].off < dec->offset) k++;
            if (k < nw->frames.size() && nw->frames[k].off == dec->offset)
            {
                nw->next = k;
                std::lock_guard<std::mutex> lock(g_win_mu);
                g_windows[dec] = nw;
                w = nw;
            } else
                delete nw;
        }
    } else
    {
        /* fresh decoder (or a position the window does not cover): new window from here */
        drop_window(dec);
        w = nullptr;
        if (fresh)
        {
            Window *nw = build_window<S>(buf, dec->offset, end_offset, window_frames());
            if (nw)
            {
                std::lock_guard<std::mutex> lock(g_win_mu);
                g_windows[dec] = nw;
                w = nw;
            }
        }
    }
    if (!w || w->next >= w->frames.size())
    {
        /* no window: one frame through the GPU frame API, state in dec->mp3d as usual */
        drop_window(dec);
        return decode_frame(&dec->mp3d, buf + dec->offset, (int)std::min<uint64_t>(end_offset - dec->offset, (uint64_t)INT_MAX), pcm, info);
    }
    const WindowFrame &f = w->frames[w->next++];
    if (f.found) *info = f.info;
    else info->frame_bytes = f.info.frame_bytes;      /* mp3dec_decode_frame leaves the other fields alone (minimp3.h:1736-1740) */
    if (f.samples)
        memcpy(pcm, (const S *)w->pcm + f.pcm_pos, (size_t)f.samples*(size_t)f.info.channels*sizeof(S));
    /* the decoder is no longer "fresh"; its carried state lives in the window, not in dec->mp3d */
    dec->mp3d.header[0] = 0xff;
    return f.samples;
}


/* ---- look-ahead for the plain frame loop ---------------------------------------------------------------------
 * The README loop -- mp3dec_init, then mp3dec_decode_frame(dec, p, left, pcm, &info); p += info.frame_bytes -- costs
 * one GPU round trip per frame when taken literally.  When a freshly initialised decoder is handed a buffer with
 * many frames, they are decoded as one window (same machinery as the streaming API) and the following calls are
 * answered from it, as long as each call continues exactly where the previous frame ended, with the same bytes and
 * the same remaining length.  Anything else -- another buffer, a shorter length, a parse-only call, the end of the
 * input -- first brings the caller's mp3dec_t up to date by replaying the last few frames (preroll_frames) from a private copy
 * of the input through the one-frame path, then continues there.  MP3B_FRAME_LOOKAHEAD=0 switches the layer off. */
struct FrameLook
{
    Window *w = nullptr;
    std::vector<uint8_t> bytes;      /* private copy of the window's input */
    const uint8_t *caller_base = nullptr;   /* caller address of bytes[0] */
    uint64_t caller_end = 0;         /* bytes from caller_base to the end of the caller's buffer */
    int float_output = 0;
    uint64_t stamp = 0;
    uint32_t id = 0;                 /* key of the window: travels in the caller's mp3dec_t as reserv = -id */
    mp3dec_t saved;                  /* w == nullptr: the decoder state the window had reached when it was retired */
    ~FrameLook() { delete w; }
};

/* The caller's mp3dec_t stays the authority on WHICH state a call continues from: while frames are served from a
 * window the struct carries the window's key in `reserv` (negative: no real decoder has that), so a copied, moved or
 * restored mp3dec_t finds its window by value, not by address, and the frame it continues with is identified by the
 * input pointer.  A key whose window is gone (evicted) restarts the decoder. */
std::mutex g_look_mu;
std::unordered_map<uint32_t, FrameLook *> g_looks;
uint64_t g_look_stamp = 0;
uint32_t g_look_next_id = 0;
constexpr size_t kMaxLooks = 16;      /* windows alive at once (~20 MB each) */
constexpr size_t kMaxSaved = 4096;    /* retired windows: saved decoder states */
constexpr int kMinLookBytes = 16384;

bool lookahead_enabled()
{
    static const bool on = [] { const char *e = getenv("MP3B_FRAME_LOOKAHEAD"); return !(e && e[0] == '0'); }();
    return on;
}

FrameLook *take_look(uint32_t id)
{
    std::lock_guard<std::mutex> lock(g_look_mu);
    auto it = g_looks.find(id);
    if (it == g_looks.end()) return nullptr;
    FrameLook *l = it->second;
    g_looks.erase(it);
    return l;
}

void replay_into(mp3dec_t *dec, const FrameLook *l);

void put_look(FrameLook *l)
{
    std::lock_guard<std::mutex> lock(g_look_mu);
    l->stamp = ++g_look_stamp;
    while (!l->id)
    {
        g_look_next_id = (g_look_next_id + 1) & 0x7fffffffu;
        if (g_look_next_id && !g_looks.count(g_look_next_id)) l->id = g_look_next_id;
    }
    size_t live = 0;
    for (auto &kv : g_looks) live += kv.second->w != nullptr;
    if (live >= kMaxLooks)
    {   /* too many decoders mid-stream: the one idle for longest keeps only its state (7 KB), not its window */
        FrameLook *old = nullptr;
        for (auto &kv : g_looks)
            if (kv.second->w && (!old || kv.second->stamp < old->stamp)) old = kv.second;
        if (old)
        {
            replay_into(&old->saved, old);
            delete old->w;
            old->w = nullptr;
            std::vector<uint8_t>().swap(old->bytes);
        }
    }
    if (g_looks.size() >= kMaxSaved)
    {   /* abandoned long ago */
        auto oldest = g_looks.begin();
        for (auto it = g_looks.begin(); it != g_looks.end(); ++it)
            if (it->second->stamp < oldest->second->stamp) oldest = it;
        delete oldest->second;
        g_looks.erase(oldest);
    }
    g_looks[l->id] = l;
}

/* bring *dec to the state it has after the frames handed out so far: fresh decoder + the last few of them */
void replay_into(mp3dec_t *dec, const FrameLook *l)
{
    const Window *w = l->w;
    mp3dec_init(dec);
    if (!w->next) return;
    const size_t first = w->next - preroll_frames(w->frames, w->next, l->bytes.data(), w->frames[0].off);
    static thread_local float scratch[MINIMP3_MAX_SAMPLES_PER_FRAME];
    const uint64_t base = w->frames[0].off;
    for (size_t k = first; k < w->next; k++)
    {
        mp3dec_frame_info_t fi;
        memset(&fi, 0, sizeof(fi));
        const uint64_t off = w->frames[k].off - base;
        mp3dec_decode_frame_direct(dec, l->bytes.data() + off, (int)std::min<uint64_t>(l->bytes.size() - off, (uint64_t)INT_MAX),
                                   scratch, &fi, l->float_output);
    }
}

template <class S>
FrameLook *start_look(const uint8_t *mp3, int mp3_bytes, const uint8_t *preroll, size_t preroll_bytes)
{
    /* window input = [frames to re-decode for the state] + [the caller's buffer]; only the extent of the frames
     * that end up in the window is kept */
    /* no frame is longer than 2881 bytes (MPEG-1 Layer II 384 kbit/s at 32 kHz; free format is capped at 2304): a
     * view of that many bytes per frame holds every frame the window can take, the rest of a long buffer stays put */
    const size_t view = std::min<size_t>((size_t)mp3_bytes, (window_frames() + kPrerollMax + 2)*2900);
    std::vector<uint8_t> in(preroll_bytes + view);
    if (preroll_bytes) memcpy(in.data(), preroll, preroll_bytes);
    memcpy(in.data() + preroll_bytes, mp3, view);
    Window *w = build_window<S>(in.data(), 0, in.size(), window_frames() + kPrerollMax);
    if (!w) return nullptr;
    size_t k = 0;
    while (k < w->frames.size() && w->frames[k].off < preroll_bytes) k++;
    if (k >= w->frames.size() || w->frames[k].off != preroll_bytes) { delete w; return nullptr; }
    w->next = k;
    FrameLook *l = new FrameLook();
    const WindowFrame &last = w->frames.back();
    in.resize((size_t)(last.off + (uint64_t)last.info.frame_bytes));
    l->bytes.swap(in);
    l->w = w;
    l->caller_base = mp3 - preroll_bytes;          /* address bytes[0] would have in the caller's buffer */
    l->caller_end = preroll_bytes + (uint64_t)mp3_bytes;
    l->float_output = sizeof(S) == 4;
    return l;
}

/* hands out the next frame of the window the way mp3dec_decode_frame reports it */
template <class S>
int serve_frame(FrameLook *l, mp3dec_t *dec, const uint8_t *mp3, S *pcm, mp3dec_frame_info_t *info)
{
    Window *w = l->w;
    const WindowFrame &f = w->frames[w->next++];
    if (f.found)
    {
        *info = f.info;
        memcpy(dec->header, mp3 + f.info.frame_offset, 4);
    } else
        info->frame_bytes = f.info.frame_bytes;      /* the other fields stay as the caller had them */
    if (f.samples) memcpy(pcm, (const S *)w->pcm + f.pcm_pos, (size_t)f.samples*(size_t)f.info.channels*sizeof(S));
    if (!dec->header[0]) dec->header[0] = 0xff;      /* "not freshly initialised": the state lives in the window */
    put_look(l);
    dec->reserv = -(int)l->id;
    return f.samples;
}

template <class S>
int frame_lookahead_t(mp3dec_t *dec, const uint8_t *mp3, int mp3_bytes, S *pcm, mp3dec_frame_info_t *info)
{
    const bool fresh = dec->header[0] == 0;
    const uint32_t id = (!fresh && dec->reserv < 0) ? (uint32_t)(-(int64_t)dec->reserv) : 0u;
    FrameLook *l = id ? take_look(id) : nullptr;
    if (id && !l)
    {
        /* the struct points at a window that no longer exists: nothing to continue from, resynchronise */
        mp3dec_init(dec);
        dec->reserv = 0;
        return -1;
    }
    if (l && !l->w)
    {
        /* this decoder's window was retired while it was idle: its state was saved, hand it back */
        memcpy(dec, &l->saved, sizeof(mp3dec_t));
        delete l;
        return -1;
    }
    if (l)
    {
        Window *w = l->w;
        /* which frame does THIS struct continue with?  Usually the window's next one; a copy or an older snapshot
         * of the struct is placed by its input pointer */
        if (mp3 >= l->caller_base && (uint64_t)(mp3 - l->caller_base) < l->caller_end)
        {
            const uint64_t want = (uint64_t)(mp3 - l->caller_base);
            if (!(w->next < w->frames.size() && w->frames[w->next].off == want))
            {
                auto it = std::lower_bound(w->frames.begin(), w->frames.end(), want,
                                           [](const WindowFrame &f, uint64_t o) { return f.off < o; });
                if (it != w->frames.end() && it->off == want) w->next = (size_t)(it - w->frames.begin());
            }
        }
        const uint64_t pos = w->next < w->frames.size() ? w->frames[w->next].off
                                                        : w->frames.back().off + (uint64_t)w->frames.back().info.frame_bytes;
        const bool here = pcm && l->float_output == (sizeof(S) == 4) && mp3 == l->caller_base + pos &&
                          (uint64_t)std::max(mp3_bytes, 0) == l->caller_end - pos;
        if (here && w->next < w->frames.size() &&
            !memcmp(mp3, l->bytes.data() + pos, (size_t)w->frames[w->next].info.frame_bytes))
            return serve_frame<S>(l, dec, mp3, pcm, info);
        if (here && w->next == w->frames.size() && mp3_bytes >= kMinLookBytes)
        {
            /* window used up, the caller goes on: next window, state from the last few frames */
            const size_t back = preroll_frames(w->frames, w->frames.size(), l->bytes.data(), w->frames[0].off);
            const uint64_t restart = w->frames[w->frames.size() - back].off;
            FrameLook *nl = start_look<S>(mp3, mp3_bytes, l->bytes.data() + restart, (size_t)(pos - restart));
            if (nl)
            {
                delete l;
                return serve_frame<S>(nl, dec, mp3, pcm, info);
            }
        }
        /* the caller left the predicted path (or the input ends): make dec current, the direct path takes over */
        replay_into(dec, l);
        delete l;
        return -1;
    }
    if (!fresh || !pcm || mp3_bytes < kMinLookBytes) return -1;
    FrameLook *nl = start_look<S>(mp3, mp3_bytes, nullptr, 0);
    if (!nl || nl->w->frames.size() - nl->w->next < 8) { delete nl; return -1; }
    return serve_frame<S>(nl, dec, mp3, pcm, info);
}

/* Decodes frames from dec->offset until dec->buffer holds deliverable samples (minimp3_ex.h:922-975, buffer mode):
 * samples still owed to a seek (to_skip) are dropped on the way, a change of sample rate / layer -- or of the channel
 * count, unless MP3D_ALLOW_MONO_STEREO_TRANSITION -- makes the error sticky.  false: end of input or error. */
template <class S>
bool refill(ExT<S> *dec, mp3dec_frame_info_t *fi)
{
    const uint64_t end = dec->end_offset ? dec->end_offset : dec->file.size;
    while (dec->buffer_consumed == dec->buffer_samples)
    {
        if (dec->offset >= end) return false;
        const uint8_t *at = dec->file.buffer + dec->offset;
        const int per_channel = window_decode_frame(dec, end, dec->buffer, fi);
        dec->buffer_consumed = 0;
        dec->buffer_samples = per_channel*fi->channels;
        const bool format_changed = dec->info.hz != fi->hz || dec->info.layer != fi->layer;
        if (format_changed)
        {
            dec->buffer_samples = per_channel;        /* as the reference leaves it */
            dec->last_error = MP3D_E_DECODE;
            return false;
        }
        if (per_channel)
        {
            const int dropped = std::min(dec->buffer_samples, std::max(dec->to_skip, 0));
            dec->buffer_consumed = dropped;
            dec->to_skip -= dropped;
            const bool channels_changed = dec->info.channels != fi->channels && !(dec->flags & MP3D_ALLOW_MONO_STEREO_TRANSITION);
            if (channels_changed && dec->buffer_consumed != dec->buffer_samples)
            {
                dec->last_error = MP3D_E_DECODE;
                return false;
            }
        } else if (dec->to_skip > 0)
            dec->to_skip -= std::min((int)hdr_frame_samples(at)*fi->channels, dec->to_skip);   /* an undecodable frame still counts */
        dec->offset += (uint64_t)fi->frame_bytes;
    }
    return true;
}

template <class S>
size_t ex_read_frame(ExT<S> *dec, S **buf, mp3dec_frame_info_t *frame_info, size_t max_samples)
{
    if (!dec || !buf || !frame_info)
    {
        if (dec) dec->last_error = MP3D_E_PARAM;
        return 0;
    }
    const bool length_known = dec->detected_samples != 0;
    if ((length_known && dec->cur_sample >= dec->detected_samples) || dec->last_error) return 0;
    *buf = NULL;
    if (!refill(dec, frame_info)) return 0;
    size_t n = std::min((size_t)(dec->buffer_samples - dec->buffer_consumed), max_samples);
    if (length_known && dec->cur_sample + n >= dec->detected_samples) n = (size_t)(dec->detected_samples - dec->cur_sample);
    *buf = dec->buffer + dec->buffer_consumed;
    dec->buffer_consumed += (int)n;
    dec->cur_sample += n;
    return n;
}

template <class S>
size_t ex_read(ExT<S> *dec, S *buf, size_t samples)
{
    if (!dec || !buf)
    {
        if (dec) dec->last_error = MP3D_E_PARAM;
        return 0;
    }
    mp3dec_frame_info_t fi;
    memset(&fi, 0, sizeof(fi));
    size_t done = 0;
    for (S *chunk = NULL; done < samples; )
    {
        const size_t n = ex_read_frame(dec, &chunk, &fi, samples - done);
        if (!n) break;
        memcpy(buf + done, chunk, n*sizeof(S));
        done += n;
    }
    return done;
}

template <class S>
void ex_close(ExT<S> *dec)
{
    if (!dec) return;
    drop_window(dec);
    if (dec->is_file && dec->file.buffer) free((void *)dec->file.buffer);
    if (dec->index.frames) free(dec->index.frames);
    memset(dec, 0, sizeof(*dec));
}

/* whole stream through the caller's I/O callbacks (seek to 0, read until short) */
int slurp_io(mp3dec_io_t *io, uint8_t **out, size_t *out_size)
{
    if (io->seek(0, io->seek_data)) return MP3D_E_IOERROR;
    size_t cap = 1 << 20, n = 0;
    uint8_t *p = (uint8_t *)malloc(cap);
    if (!p) return MP3D_E_MEMORY;
    for (;;)
    {
        if (n == cap)
        {
            uint8_t *g = (uint8_t *)realloc(p, cap*2);
            if (!g) { free(p); return MP3D_E_MEMORY; }
            p = g; cap *= 2;
        }
        const size_t want = cap - n;
        size_t r = io->read(p + n, want, io->read_data);
        if (r > want) { free(p); return MP3D_E_IOERROR; }
        n += r;
        if (r != want) break;     /* short read: end of stream */
    }
    *out = p; *out_size = n;
    return 0;
}

int slurp_file(const char *name, uint8_t **out, size_t *out_size)
{
    if (!name) return MP3D_E_PARAM;
    FILE *f = fopen(name, "rb");
    if (!f) return MP3D_E_IOERROR;
    int res = MP3D_E_IOERROR;
    uint8_t *p = NULL;
    long size = -1;
    if (!fseek(f, 0, SEEK_END) && (size = ftell(f)) >= 0 && !fseek(f, 0, SEEK_SET))
    {
        p = (uint8_t *)malloc(size ? (size_t)size : 1);
        if (!p) res = MP3D_E_MEMORY;
        else if (fread(p, 1, (size_t)size, f) == (size_t)size) res = 0;
    }
    fclose(f);
    if (res) { free(p); return res; }
    *out = p; *out_size = (size_t)size;
    return 0;
}

template <class S>
int ex_open_owned(ExT<S> *dec, uint8_t *p, size_t n, int flags)
{
    int ret = ex_open_buf(dec, p, n, flags);
    if (ret) { free(p); return ret; }
    dec->is_file = 1;
    return 0;
}

/* mp3dec_load_buf with a progress callback (minimp3_ex.h:313-522 restated): the callback sees every decoded
 * frame in order and may cancel, so this flavour walks the stream through the single-frame GPU path instead of the
 * one-shot batch. */
template <class S>
int load_buf_frames(mp3dec_t *dec, const uint8_t *buf, size_t buf_size, mp3dec_file_info_t *info,
                    MP3D_PROGRESS_CB progress_cb, void *user_data)
{
    const size_t orig_size = buf_size;
    memset(info, 0, sizeof(*info));
    mp3dec_init(dec);
    StreamHead head;
    if (!open_stream_head(&buf, &buf_size, &head)) return 0;
    mp3dec_frame_info_t fi;
    memset(&fi, 0, sizeof(fi));
    fi.channels = head.first.channels; fi.hz = head.first.hz; fi.layer = head.first.layer;
    fi.bitrate_kbps = head.first.bitrate_kbps; fi.frame_bytes = head.first.frame_bytes;
    const uint64_t detected_samples = head.detected_samples;
    size_t to_skip = (size_t)head.to_skip;
    const int frame_samples = head.frame_samples*fi.channels;
    size_t allocated = MINIMP3_MAX_SAMPLES_PER_FRAME*sizeof(S);
    if (detected_samples) allocated += (size_t)detected_samples*sizeof(S);
    else allocated += (buf_size/fi.frame_bytes)*(size_t)frame_samples*sizeof(S);
    S *out = (S *)malloc(allocated);
    if (!out) return MP3D_E_MEMORY;
    info->channels = fi.channels;
    info->hz = fi.hz;
    info->layer = fi.layer;
    size_t n_samples = 0, avg_kbps = 0, frames = 0;
    int ret = 0;
    /* the frames come out of look-ahead windows (one GPU batch per 4096 frames), the callback still sees them one
     * by one: a cancelled load has merely decoded a little ahead */
    ExT<S> *reader = (ExT<S> *)calloc(1, sizeof(ExT<S>));
    if (!reader) { free(out); return MP3D_E_MEMORY; }
    reader->file.buffer = buf;
    reader->file.size = buf_size;
    const uint64_t stream_end = buf_size;
    do
    {
        if (allocated - n_samples*sizeof(S) < MINIMP3_MAX_SAMPLES_PER_FRAME*sizeof(S))
        {
            allocated *= 2;
            S *grown = (S *)realloc(out, allocated);
            if (!grown) { free(out); drop_window(reader); free(reader); return MP3D_E_MEMORY; }
            out = grown;
        }
        size_t samples = (size_t)window_decode_frame(reader, stream_end, out + n_samples, &fi);
        reader->offset += (uint64_t)fi.frame_bytes;
        buf_size -= fi.frame_bytes;
        if (samples)
        {
            if (info->hz != fi.hz || info->layer != fi.layer) { ret = MP3D_E_DECODE; break; }
            if (info->channels && info->channels != fi.channels) { ret = MP3D_E_DECODE; break; }
            samples *= fi.channels;
            if (to_skip)
            {
                const size_t skip = std::min(samples, to_skip);
                to_skip -= skip;
                samples -= skip;
                memmove(out, out + skip, samples*sizeof(S));
            }
            n_samples += samples;
            avg_kbps += fi.bitrate_kbps;
            frames++;
            ret = progress_cb(user_data, orig_size, orig_size - buf_size, &fi);
            if (ret) break;
        }
    } while (fi.frame_bytes);
    drop_window(reader);
    free(reader);
    (void)dec;
    if (detected_samples && n_samples > detected_samples) n_samples = (size_t)detected_samples;
    if (allocated != n_samples*sizeof(S))
    {
        S *fit = (S *)realloc(out, n_samples*sizeof(S));
        if (!fit && n_samples) { free(out); return MP3D_E_MEMORY; }
        out = fit;
    }
    info->buffer = (mp3d_sample_t *)out;
    info->samples = n_samples;
    if (frames) info->avg_bitrate_kbps = (int)(avg_kbps/frames);
    return ret;
}

}  // namespace

extern "C" __attribute__((visibility("hidden"))) int mp3dec_frame_lookahead(mp3dec_t *dec, const uint8_t *mp3, int mp3_bytes, void *pcm,
                                                                            mp3dec_frame_info_t *info, int float_output)
{
    if (!lookahead_enabled() || !dec || !mp3 || !info) return -1;
    return float_output ? frame_lookahead_t<float>(dec, mp3, mp3_bytes, (float *)pcm, info)
                        : frame_lookahead_t<int16_t>(dec, mp3, mp3_bytes, (int16_t *)pcm, info);
}

extern "C" {

int mp3dec_iterate_buf(const uint8_t *buf, size_t buf_size, MP3D_ITERATE_CB callback, void *user_data)
{
    if (!buf || (size_t)-1 == buf_size || !callback) return MP3D_E_PARAM;
    return walk_frames(buf, buf_size, [&](const uint8_t *frame, int frame_size, int free_format_bytes, size_t left, uint64_t offset,
                                          mp3dec_frame_info_t *fi) {
        return callback(user_data, frame, frame_size, free_format_bytes, left, offset, fi);
    });
}

int mp3dec_iterate_cb(mp3dec_io_t *io, uint8_t *buf, size_t buf_size, MP3D_ITERATE_CB callback, void *user_data)
{
    if (!io || !buf || (size_t)-1 == buf_size || buf_size < MINIMP3_BUF_SIZE || !callback) return MP3D_E_PARAM;
    uint8_t *p; size_t n;
    int r = slurp_io(io, &p, &n);
    if (r) return r;
    r = mp3dec_iterate_buf(p, n, callback, user_data);
    free(p);
    return r;
}

int mp3dec_detect_cb(mp3dec_io_t *io, uint8_t *buf, size_t buf_size)
{
    if (!buf || (size_t)-1 == buf_size || (io && buf_size < MINIMP3_BUF_SIZE)) return MP3D_E_PARAM;
    if (!io) return mp3dec_detect_buf(buf, buf_size);
    uint8_t *p; size_t n;
    int r = slurp_io(io, &p, &n);
    if (r) return r;
    r = mp3dec_detect_buf(p, n);
    free(p);
    return r;
}

__attribute__((visibility("hidden"))) int mp3dec_load_buf_progress(mp3dec_t *dec, const uint8_t *buf, size_t buf_size, mp3dec_file_info_t *info,
                                        MP3D_PROGRESS_CB progress_cb, void *user_data, int float_output)
{
    return float_output ? load_buf_frames<float>(dec, buf, buf_size, info, progress_cb, user_data)
                        : load_buf_frames<int16_t>(dec, buf, buf_size, info, progress_cb, user_data);
}

#define LOAD_CB_IMPL(NAME, LOADBUF)                                                                                  \
    int NAME(mp3dec_t *dec, mp3dec_io_t *io, uint8_t *buf, size_t buf_size, mp3dec_file_info_t *info,                \
             MP3D_PROGRESS_CB progress_cb, void *user_data)                                                          \
    {                                                                                                                \
        if (!dec || !buf || !info || (size_t)-1 == buf_size || (io && buf_size < MINIMP3_BUF_SIZE)) return MP3D_E_PARAM; \
        if (!io) return LOADBUF(dec, buf, buf_size, info, progress_cb, user_data);                                   \
        uint8_t *p; size_t n;                                                                                        \
        int r = slurp_io(io, &p, &n);                                                                                \
        if (r) return r;                                                                                             \
        r = LOADBUF(dec, p, n, info, progress_cb, user_data);                                                        \
        free(p);                                                                                                     \
        return r;                                                                                                    \
    }
int mp3dec_load_buf_f32(mp3dec_t *, const uint8_t *, size_t, mp3dec_file_info_t *, MP3D_PROGRESS_CB, void *);
LOAD_CB_IMPL(mp3dec_load_cb, mp3dec_load_buf)
LOAD_CB_IMPL(mp3dec_load_cb_f32, mp3dec_load_buf_f32)

#define EX_API(SUFFIX, S)                                                                                            \
    int mp3dec_ex_open_buf##SUFFIX(mp3dec_ex_t *dec, const uint8_t *buf, size_t buf_size, int flags)                 \
    { return ex_open_buf((ExT<S> *)dec, buf, buf_size, flags); }                                                     \
    int mp3dec_ex_open_cb##SUFFIX(mp3dec_ex_t *dec, mp3dec_io_t *io, int flags)                                      \
    {                                                                                                                \
        if (!dec || !io || (flags & (~MP3D_FLAGS_MASK))) return MP3D_E_PARAM;                                        \
        uint8_t *p; size_t n;                                                                                        \
        int r = slurp_io(io, &p, &n);                                                                                \
        if (r) return r;                                                                                             \
        return ex_open_owned((ExT<S> *)dec, p, n, flags);                                                            \
    }                                                                                                                \
    int mp3dec_ex_open##SUFFIX(mp3dec_ex_t *dec, const char *file_name, int flags)                                   \
    {                                                                                                                \
        if (!dec) return MP3D_E_PARAM;                                                                               \
        uint8_t *p; size_t n;                                                                                        \
        int r = slurp_file(file_name, &p, &n);                                                                       \
        if (r) return r;                                                                                             \
        return ex_open_owned((ExT<S> *)dec, p, n, flags);                                                            \
    }                                                                                                                \
    void mp3dec_ex_close##SUFFIX(mp3dec_ex_t *dec) { ex_close((ExT<S> *)dec); }                                      \
    int mp3dec_ex_seek##SUFFIX(mp3dec_ex_t *dec, uint64_t position) { return ex_seek((ExT<S> *)dec, position); }     \
    size_t mp3dec_ex_read_frame##SUFFIX(mp3dec_ex_t *dec, S **buf, mp3dec_frame_info_t *fi, size_t max_samples)      \
    { return ex_read_frame((ExT<S> *)dec, buf, fi, max_samples); }                                                   \
    size_t mp3dec_ex_read##SUFFIX(mp3dec_ex_t *dec, S *buf, size_t samples) { return ex_read((ExT<S> *)dec, buf, samples); }

EX_API(, int16_t)
EX_API(_f32, float)

int mp3dec_detect(const char *file_name)
{
    uint8_t *p; size_t n;
    int r = slurp_file(file_name, &p, &n);
    if (r) return r;
    r = mp3dec_detect_buf(p, n);
    free(p);
    return r;
}

int mp3dec_iterate(const char *file_name, MP3D_ITERATE_CB callback, void *user_data)
{
    uint8_t *p; size_t n;
    int r = slurp_file(file_name, &p, &n);
    if (r) return r;
    r = mp3dec_iterate_buf(p, n, callback, user_data);
    free(p);
    return r;
}

#define LOAD_FILE_IMPL(NAME, LOADBUF)                                                                                \
    int NAME(mp3dec_t *dec, const char *file_name, mp3dec_file_info_t *info, MP3D_PROGRESS_CB progress_cb, void *user_data) \
    {                                                                                                                \
        uint8_t *p; size_t n;                                                                                        \
        int r = slurp_file(file_name, &p, &n);                                                                       \
        if (r) return r;                                                                                             \
        r = LOADBUF(dec, p, n, info, progress_cb, user_data);                                                        \
        free(p);                                                                                                     \
        return r;                                                                                                    \
    }
LOAD_FILE_IMPL(mp3dec_load, mp3dec_load_buf)
LOAD_FILE_IMPL(mp3dec_load_f32, mp3dec_load_buf_f32)

}  // extern "C"
