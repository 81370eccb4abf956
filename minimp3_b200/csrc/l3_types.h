/*
 * l3_types.h -- records shared by the host bitstream stage and the CUDA kernels.
 *
 * Data layout in HBM:
 *   main data   uint8  [total_bytes]      per stream: the frame payloads that follow the side info,
 *                                         concatenated ("C" in SURVEY.md 8a-r2), 16-byte aligned start
 *   GcDesc      32 B   [n_gc]             one per granule-channel, order = stream, frame, granule, channel
 *   TileDesc    32 B   [n_tiles]          one per run of <= MP3B_TILE_GRANULES consecutive granules
 *   xr          float  [n_gc][576]        requantised (+ MS stereo) spectra, written by the entropy kernel
 *   pcm         int16 / float [n_samples] interleaved PCM, stream after stream
 */
#ifndef MP3B_L3_TYPES_H
#define MP3B_L3_TYPES_H
#include <stdint.h>

#define MP3B_TILE_GRANULES 8
/* mono Layer III tiles hold twice the granules: the transform CTA's shared memory is laid out for two channels */
#define MP3B_TILE_GRANULES_MONO 16

/* GcDesc.flags0 */
#define GCF_BLOCK_TYPE_MASK 0x03
#define GCF_MIXED           0x04
#define GCF_PREFLAG         0x08
#define GCF_SCALEFAC_SCALE  0x10
#define GCF_COUNT1_TABLE    0x20
#define GCF_MS_STEREO       0x40   /* header: joint stereo with MS bit (HDR_IS_MS_STEREO / HDR_TEST_MS_STEREO) */
#define GCF_I_STEREO        0x80   /* header: intensity stereo bit (HDR_TEST_I_STEREO) */
/* GcDesc.flags1 */
#define GCF1_MPEG1          0x01
#define GCF1_CH             0x02   /* this is channel 1 */
#define GCF1_STEREO         0x04   /* frame has two channels */
#define GCF1_GR1            0x08   /* second granule of an MPEG-1 frame (scfsi source = gc - nch) */
#define GCF1_MS_PLAIN       0x10   /* MS stereo handled by the entropy kernel (MS bit set, intensity bit clear) */
#define GCF1_MS_TEST        0x20   /* raw MS bit of the header (HDR_TEST_MS_STEREO), whatever the channel mode */

/* One granule-channel: the L3_gr_info_t of the reference (minimp3.h:223-230) with the sfb table pointer
 * replaced by (sfb kind, sample-rate class) and the bit-reservoir position resolved to an absolute offset. */
typedef struct
{
    uint64_t bit_off;            /* absolute bit offset of part 2 (scalefactors) in the main-data buffer */
    uint16_t part_23_length;
    uint16_t big_values;
    uint16_t scalefac_compress;
    uint8_t  global_gain;
    uint8_t  flags0;
    uint8_t  table_select[3];
    uint8_t  region_count[3];
    uint8_t  subblock_gain[3];
    uint8_t  scfsi;
    uint8_t  n_long_sfb, n_short_sfb;
    uint8_t  sr_idx;             /* row of the sfb tables, 0..7 */
    uint8_t  flags1;
    uint8_t  pad[2];
} GcDesc;

/* One tile of the transform kernel: n_gr consecutive decoded granules of one stream that share the
 * channel count and have no decoder reset between them.  halo[ch][0] / halo[ch][1] are the granule-channel
 * indices of the previous / second previous granule of that channel since the last reset (-1: zero state).
 * They replace the carried mdct_overlap / qmf_state of the reference (minimp3.h:18-23). */
typedef struct
{
    uint32_t gc_first;           /* index of granule 0 channel 0 */
    uint8_t  n_gr, nch, pad[2];
    uint64_t pcm_off;            /* sample offset (channel-inclusive) of the tile's first PCM sample */
    int32_t  halo[2][2];
} TileDesc;

/* What one transform CTA needs to know about its tile, laid out by k_l3_tile_meta once the extents (nzl) are
 * known so that the CTA gets everything in one round trip: per slot (granule incl. two halo granules x channel)
 * the row to copy and how to treat it, the prefix sums of the IMDCT job list, and the tile header. */
#define MP3B_TILE_SLOTS ((MP3B_TILE_GRANULES + 2)*2)
/* TileSlot.info */
#define TSI_NZL32_MASK   0x1fu           /* bits 0..4: decoded extent in units of 32 lines (0..18) */
#define TSI_ACT_SHIFT    5               /* bits 5..10: subbands that can hold non-zero lines after antialias */
#define TSI_BT_SHIFT     11              /* bits 11..12: block type */
#define TSI_MS_SHIFT     13              /* bits 13..14: 0 none, 1 row = L + R, 2 row = L - R */
#define TSI_NLONG_SHIFT  15              /* bits 15..17: leading long subbands of a mixed block (0, 2, 4) */
#define TSI_MAP_SHIFT    18              /* bits 18..22: reorder map row + 1, 0 = none */
#define TSI_PAIR_SHIFT   23              /* bits 23..27: plain-MS pair with both rows in the tile: max extent / 32, else 0 */
#define TSI_FAR_SHIFT    28              /* bit 28: plain-MS row whose partner row is not part of the tile */
/* IMDCT job list of a tile: [MP3B_TILE_JOBS] entries (slot << 8 | subband), the first TileRec.n_jobs valid:
 * only subbands that are active in the granule itself or in the one before it (whose tail overlaps into it)
 * need work, everything else is exactly zero. */
#define MP3B_TILE_JOBS (MP3B_TILE_SLOTS*32)
typedef struct
{
    int32_t  gc;                 /* granule-channel row or -1 (zero state) */
    uint32_t info;
} TileSlot;
typedef struct
{
    uint8_t  n_gr, nch;                   /* header: one 16-byte load */
    uint16_t n_jobs;                      /* entries of the tile's IMDCT job list */
    uint8_t  pad[4];
    uint64_t pcm_off;
    TileSlot slot[MP3B_TILE_SLOTS];
    uint8_t  pad2[256 - MP3B_TILE_SLOTS*8 - 16];
} TileRec;

/* One decodable Layer III frame as the walker hands it to k_l3_side_info: the raw side-info bytes and where its
 * descriptors and its main data go.  Granule 0 of the frame owns descriptors gc_first .. gc_first+nch-1,
 * granule 1 (MPEG-1) the next nch. */
typedef struct
{
    uint64_t md_bit;             /* absolute bit offset of the frame's main data in the main-data buffer */
    uint32_t gc_first;
    uint8_t  hdr[4];
    uint8_t  nch, n_gr, pad[6];
    uint8_t  si[32];             /* side info as it follows the header (and the optional CRC) */
    uint8_t  pad2[8];
} FrameRec;


/* One Layer I / II frame (minimp3.h:1783-1802): payload position + header; its output is n_blocks blocks of
 * 12 time slots x 32 subbands per channel in the subband buffer sub12[block-channel][12][32]. */
typedef struct
{
    uint64_t bit_off;            /* payload (after header and optional CRC) in the main-data buffer */
    uint32_t block_first;        /* index of block 0 channel 0 in sub12 */
    uint8_t  hdr[4];
    uint16_t kbps;               /* header bitrate, 0 = free format (selects the Layer II allocation table) */
    uint8_t  nch, n_blocks;      /* n_blocks: 1 (Layer I, 384 samples) or 3 (Layer II, 1152 samples) */
    uint8_t  pad[12];
} L12FrameDesc;

#endif
