/*
 * host_parse.h -- host side of the Layer III hot path (stays on the CPU):
 * frame sync, header math, side-info parse, bit-reservoir resolution and tile planning.
 *
 * This restates, as a frame walker that never touches PCM, the control flow of
 *   mp3dec_decode_frame      /root/reference/minimp3.h:1713-1806
 *   mp3d_find_frame/match    /root/reference/minimp3.h:1657-1706
 *   L3_read_side_info        /root/reference/minimp3.h:484-607
 *   L3_restore/save_reservoir/root/reference/minimp3.h:1212-1236
 * and, on top of it, the whole-file semantics of mp3dec_load_buf (minimp3_ex.h:308-525).
 */
#ifndef MP3B_HOST_PARSE_H
#define MP3B_HOST_PARSE_H
#include <stddef.h>
#include <stdint.h>

#include <vector>

#include "l3_types.h"
#include "l3_sideinfo.cuh"

namespace mp3b {

struct FrameInfo
{
    int frame_bytes, frame_offset, channels, hz, layer, bitrate_kbps;
};

/* header helpers (minimp3.h:61-78, 264-314) */
bool hdr_valid(const uint8_t *h);
bool hdr_compare(const uint8_t *h1, const uint8_t *h2);
unsigned hdr_bitrate_kbps(const uint8_t *h);
unsigned hdr_sample_rate_hz(const uint8_t *h);
unsigned hdr_frame_samples(const uint8_t *h);
int hdr_frame_bytes(const uint8_t *h, int free_format_size);
int hdr_padding(const uint8_t *h);
inline bool hdr_is_mono(const uint8_t *h) { return (h[3] & 0xC0) == 0xC0; }
inline bool hdr_is_crc(const uint8_t *h) { return !(h[1] & 1); }
inline bool hdr_is_mpeg1(const uint8_t *h) { return (h[1] & 0x8) != 0; }
inline int hdr_layer(const uint8_t *h) { return 4 - ((h[1] >> 1) & 3); }

int find_frame(const uint8_t *mp3, int mp3_bytes, int *free_format_bytes, int *ptr_frame_bytes);

/* MSB-first bit reader over a byte buffer; reads past `limit` return 0 (minimp3.h:248-262) */
struct BitReader
{
    const uint8_t *buf;
    int pos, limit;
    void init(const uint8_t *data, int bytes) { buf = data; pos = 0; limit = bytes*8; }
    uint32_t get_slow(int start, int n) const;
    /* 1 <= n <= 32 bits, MSB first; 0 once the frame is exhausted (minimp3.h:248-262) */
    inline uint32_t get(int n)
    {
        const int start = pos;
        pos += n;
        if (pos > limit || n == 0) return 0;
        const int byte0 = start >> 3;
        if (byte0 + 8 <= (limit >> 3))
        {
            /* one unaligned 8-byte load inside the frame: a field fits behind any start bit */
            uint64_t w;
            __builtin_memcpy(&w, buf + byte0, 8);
            w = __builtin_bswap64(w);
            return (uint32_t)((w << (start & 7)) >> (64 - n));
        }
        return get_slow(start, n);
    }
};

/* returns main_data_begin or -1 (minimp3.h:484-607) */
int read_side_info(BitReader *bs, GranuleSide *gr, const uint8_t *hdr);
/* same decision and bit position, only part_23_length / big_values / block_type filled in (l3_sideinfo.cuh) */
int read_side_info_lite(BitReader *bs, GranuleSide *gr, const uint8_t *hdr);
int sample_rate_class(const uint8_t *hdr);   /* 0..7, row of the sfb tables */

/* A decoded granule of the stream in output order. */
struct GranuleRef
{
    uint32_t gc;       /* stream-relative index of channel 0's GcDesc (Layer III) or block-channel (Layer I/II) */
    uint8_t nch;
    uint8_t reset;     /* decoder state was zeroed before this granule */
    uint8_t kind;      /* 0: Layer III granule (576 samples/channel), 1: Layer I/II block (384 samples/channel) */
};

/* Everything the walker learns about one stream.  Indices are stream-relative until rebase(). */
struct StreamPlan
{
    /* results mirrored into mp3dec_file_info_t */
    int ret = 0;                 /* 0 or MP3D_E_* */
    uint64_t samples_decoded = 0;/* channel-inclusive samples the kernels produce for this stream */
    uint64_t out_skip = 0;       /* leading samples to drop (encoder delay, mp3dec_load_buf to_skip) */
    uint64_t out_samples = 0;    /* samples delivered to the caller after skip/cut */
    int channels = 0, hz = 0, layer = 0, avg_bitrate_kbps = 0;
    /* device work */
    std::vector<GcDesc> gcs;               /* host-side descriptors (single-frame API, CPU harness) ... */
    std::vector<FrameRec> frame_recs;      /* ... or, with WalkOptions.device_side_info, the frames k_l3_side_info turns into them */
    std::vector<uint32_t> gc_key;          /* device_side_info: big_values << 16 | part_23_length per descriptor slot, ~0 = padding slot */
    uint32_t n_gc = 0;                     /* descriptor slots of the stream (= gcs.size() in host mode) */
    std::vector<GranuleRef> granules;
    std::vector<TileDesc> tiles;
    std::vector<L12FrameDesc> l12_frames;  /* Layer I/II frames; block indices stream-relative */
    std::vector<TileDesc> tiles12;         /* synthesis tiles over Layer I/II blocks */
    uint32_t n_blockch = 0;                /* Layer I/II block-channels */
    uint64_t main_data_bytes = 0;
    uint32_t frames = 0, frames_decoded = 0;
    bool has_layer12 = false;
    std::vector<uint32_t> istereo_gcs;  /* stream-relative gc index (channel 0) of granules needing intensity stereo */
    std::vector<uint32_t> order;        /* real granule-channels sorted by part_23_length: the entropy kernel's lane order */
};

struct WalkOptions
{
    bool raw_frames = false;            /* plain mp3dec_decode_frame loop: no tag skipping / VBR-tag trimming */
    bool allow_mono_stereo = false;     /* MINIMP3_ALLOW_MONO_STEREO_TRANSITION */
    int tile_granules = MP3B_TILE_GRANULES;
    bool device_side_info = false;      /* emit FrameRec records instead of GcDesc: the GPU parses the side info */
};

/*
 * Walks one stream.  Main data (the concatenated payloads) is written to md_dst (capacity >= size + 16);
 * GcDesc.bit_off = md_base_bits + offset inside md_dst.
 */
void walk_stream(const uint8_t *buf, size_t size, uint8_t *md_dst, uint64_t md_base_bits,
                 const WalkOptions &opt, StreamPlan *plan);

/* tag helpers (minimp3_ex.h:137-191) */
void skip_id3(const uint8_t **pbuf, size_t *pbuf_size);
size_t skip_id3v2(const uint8_t *buf, size_t buf_size);
void skip_id3v1(const uint8_t *buf, size_t *pbuf_size);
/* minimp3_ex.h:193-263 */
int check_vbrtag(const uint8_t *frame, int frame_size, uint32_t *frames, int *delay, int *padding);

/* Head of a stream with mp3dec_load_buf semantics (minimp3_ex.h:313-371): tags skipped, first frame found, a
 * Xing / Info tag frame evaluated and dropped.  On return *pbuf / *psize stand at the first audio frame.  false:
 * nothing to decode (no frame, or a tag announcing zero samples). */
struct StreamHead
{
    FrameInfo first{};           /* parameters of the first frame (frame_bytes = its size) */
    int frame_samples = 0;       /* samples per channel of that frame */
    uint64_t to_skip = 0;        /* channel-inclusive samples of encoder delay to drop */
    uint64_t detected_samples = 0;   /* stream length the tag announces, 0 = unknown */
};
bool open_stream_head(const uint8_t **pbuf, size_t *psize, StreamHead *head);

/* mp3dec_detect_buf (minimp3_ex.h:265-306, buffer flavour) */
int detect_buf(const uint8_t *buf, size_t buf_size);

/*
 * Incremental emulation of the decoder's bitstream-side state (header, free_format_bytes, reservoir fill)
 * used by both walk_stream and the single-frame API.
 */
struct SyncState
{
    uint8_t header[4] = {0, 0, 0, 0};
    int free_format_bytes = 0;
    int reserv = 0;
};

struct FrameParse
{
    FrameInfo info{};
    int samples = 0;            /* per channel; 0 when undecodable */
    bool reset = false;         /* decoder state was cleared (memset) during this call */
    bool init_only = false;     /* decoder was re-initialised (header[0]=0) during this call */
    int layer = 0;
    /* Layer III only */
    int n_gr = 0, nch = 0;
    int main_data_begin = 0;
    const uint8_t *payload = nullptr;  /* bytes after the side info */
    int payload_bytes = 0;
    bool success = false;       /* reservoir holds main_data_begin bytes (Layer III) / payload not overrun (I/II) */
    GranuleSide side[4];
    const uint8_t *hdr = nullptr;
    const uint8_t *side_info = nullptr; /* first byte of the side info (behind header and CRC) */
};

/* One mp3dec_decode_frame call worth of parsing (no PCM).  Updates st->header/free_format_bytes and, for
 * Layer III, st->reserv exactly as L3_restore_reservoir / L3_save_reservoir would. */
void parse_frame(SyncState *st, const uint8_t *mp3, int mp3_bytes, FrameParse *out, bool headers_only = false,
                 bool lite_side_info = false);

void fill_gc_desc(GcDesc *d, const GranuleSide &g, const uint8_t *hdr, int ch, int nch, int igr, uint64_t bit_off);

}  // namespace mp3b
#endif
