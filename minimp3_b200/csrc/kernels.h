/*
 * kernels.h -- launch interface of the sm_100a kernels (internal; the C ABI is include/minimp3_b200.h).
 */
#ifndef MP3B_KERNELS_H
#define MP3B_KERNELS_H
#include <cuda_runtime.h>
#include <stdint.h>

#include "l3_types.h"

namespace mp3b {

/* Device-resident constant tables (uploaded once per device by tables_upload). */
struct DeviceTables
{
    uint16_t *reorder_map;   /* [16][576]: (sr_idx*2 + mixed) -> source line of short-block destination line e */
    float *fir_coef;         /* [32][16]: per output phase j the 16 signed window taps */
    uint8_t *band_of_line;   /* [3][8][576]: (kind: 0 long, 1 mixed, 2 short; sr_idx) -> scalefactor band of spectral line */
};

struct EntropyArgs
{
    const uint32_t *md_words;     /* main data, viewed as 32-bit words (byte order fixed in the kernel) */
    uint64_t md_n_words;
    const GcDesc *gcs;
    uint32_t n_gc;                /* multiple of 2; stereo pairs start at even indices */
    const uint32_t *order;        /* lane order: granule-channel index handled by work item k */
    uint32_t n_order;
    uint32_t *group_counter;      /* scratch word: launch_entropy clears it, the warps claim their next 32 work items from it */
    float *xr;                    /* [n_gc][576] */
    uint16_t *nzl;                /* [n_gc] lines written per row (multiple of 32) */
    uint8_t *ist_pos;             /* [n_gc][40] written for intensity-stereo frames (and in tap mode) */
    /* optional taps for the parity tests (null in production) */
    int16_t *tap_q;               /* [n_gc][576] */
    float *tap_scf;               /* [n_gc][40]  */
};

struct StereoArgs
{
    const GcDesc *gcs;
    const uint32_t *granules;     /* gc index of channel 0 of every granule that needs intensity stereo */
    uint32_t n_granules;
    float *xr;
    uint16_t *nzl;
    const uint8_t *ist_pos;
};

struct TransformArgs
{
    const GcDesc *gcs;
    const TileDesc *tiles;
    uint32_t n_tiles;
    const float *xr;
    const uint16_t *nzl;
    TileRec *tile_recs;           /* scratch [n_tiles], written by k_l3_tile_meta, read by k_l3_transform */
    uint16_t *tile_jobs;          /* scratch [n_tiles][MP3B_TILE_JOBS], likewise */
    void *pcm;                    /* int16 or float, interleaved */
    int float_output;
    /* optional carried state for the single-frame API (null in batch mode): mp3dec_t.mdct_overlap / qmf_state */
    const float *state_in;        /* [2][288] tails + [2][15][32] synthesis history */
    float *state_out;
    /* optional tap (tests): subband samples after IMDCT + sign, [n_gc][576] */
    float *tap_sub;
    /* optional profiling aid: clock64() of thread 0 at every phase boundary, [n_tiles][16] */
    long long *phase_clk;
};

struct L12Compact;     /* l12_decode.cuh */

struct Layer12Args
{
    const uint32_t *md_words;
    uint64_t md_n_words;
    const L12FrameDesc *frames;
    uint32_t n_frames;
    float *sub12;                 /* [n_blockch][12][32] scaled subband samples */
    L12Compact *info;             /* scratch [n_frames]: the parsed scale info, written by k_l12_scale_info */
};

/* synthesis over Layer I/II blocks: TileDesc.gc_first / halo index block-channels of sub12 */
struct Synth12Args
{
    const TileDesc *tiles;
    uint32_t n_tiles;
    const float *sub12;
    void *pcm;
    int float_output;
    const float *state_in;        /* optional carried history [2][15][32] (single-frame API) */
    float *state_out;
};

/* ---- GPU frame scan (k_scan.cu) ---- */
#define SCAN_OK        0
#define SCAN_IRREGULAR 1      /* the host walker takes the stream (resync, channel change, bad side info) */
#define SCAN_TAIL      2      /* regular up to tail_pos; what follows is too short for a frame or not a frame (<= SCAN_TAIL_MAX bytes) */
#define SCAN_TAIL_MAX  16384u
#define SCANF_DECODABLE 1     /* the bit reservoir holds the frame's main_data_begin bytes */
#define SCANF_ISTEREO   2     /* stereo frame with the intensity bit set */

struct ScanStream              /* host -> device, one per stream handed to the scan */
{
    uint64_t src_off;          /* first audio frame of the stream in the raw buffer */
    uint32_t size;             /* bytes from there to the end of the stream */
    uint32_t slot_base;        /* first ScanFrame slot */
    uint32_t slot_cap;         /* slots the stream may fill: size / shortest possible frame */
    uint8_t first_hdr[4];      /* header of the first frame (found and verified by the host's frame search) */
};

struct ScanFrame               /* one frame as the scan found it */
{
    uint32_t pos;              /* offset of the header from src_off */
    uint16_t frame_size;
    uint8_t payload_off;       /* header + CRC + side info */
    uint8_t flags;             /* SCANF_* */
    uint16_t main_data_begin;
    uint16_t pad;
    uint32_t md_before;        /* main-data bytes of the stream in front of this frame's payload */
    uint32_t dec_index;        /* decodable frames in front of this one */
    uint32_t ist_before;       /* intensity-stereo granules in front of this one */
};

struct ScanSummary
{
    uint32_t n_frames, n_dec, n_ist, kbps_sum, md_len;
    uint8_t status, nch, n_gr, pad;
    uint32_t tail_pos;         /* where the scan stopped (offset from the first audio frame) */
    uint32_t pad2;
};

struct StreamBase              /* host -> device after the prefix sums: where a stream's records go */
{
    uint64_t md_base, pcm_base;
    uint32_t n_frames, n_dec;
    uint32_t gc_base, gc_pad_before, fr_base, tile_base, ist_base;
    uint32_t frame_base;       /* frames (decodable or not) of the taken streams in front of this one; monotonic over all streams */
    uint8_t take, nch, n_gr, pad;
};

struct ScanArgs
{
    const uint8_t *raw;
    const ScanStream *streams;
    uint32_t n_streams;
    ScanFrame *frames;
    ScanSummary *summaries;
};

struct EmitArgs
{
    const uint8_t *raw;
    const ScanStream *streams;
    const StreamBase *bases;
    uint32_t n_streams;
    const ScanFrame *frames;
    uint32_t n_frames;             /* frames of all taken streams */
    FrameRec *frame_recs;
    uint32_t *order;
    uint32_t *ist_list;
    uint8_t *md;
    TileDesc *tiles;
    uint32_t n_tiles;
};

cudaError_t launch_stream_scan(const ScanArgs &a, cudaStream_t s);
cudaError_t launch_stream_emit(const EmitArgs &a, cudaStream_t s);

int tables_upload(DeviceTables *t);
void tables_free(DeviceTables *t);

cudaError_t launch_entropy(const DeviceTables &t, const EntropyArgs &a, cudaStream_t s);
cudaError_t launch_intensity_stereo(const DeviceTables &t, const StereoArgs &a, cudaStream_t s);
/* Layer III side info of n_frames frames -> their GcDesc records (k_sideinfo.cu) */
cudaError_t launch_side_info(const FrameRec *frames, uint32_t n_frames, GcDesc *gcs, cudaStream_t s);
/* Lane order of the entropy kernel across the streams of a batch (k_order.cu): order_in[0..n) -> order_out, sorted
 * inside windows of 32 streams by (big_values, part_23_length / 8), equal keys position by position -- where the
 * streams of a window are alike; other windows keep the walker's per-stream order.  gc_base[i]:
 * first descriptor slot of stream i (n_streams + 1 entries), max_slots: most slots any stream has. */
size_t order_sort_scratch_bytes(uint32_t n, uint32_t n_streams);
cudaError_t launch_order_sort(const GcDesc *gcs, const uint32_t *order_in, uint32_t n, uint32_t *order_out,
                              const uint32_t *gc_base, uint32_t n_streams, uint32_t max_slots,
                              void *scratch, size_t scratch_bytes, cudaStream_t s);
cudaError_t launch_transform(const DeviceTables &t, const TransformArgs &a, cudaStream_t s);
cudaError_t launch_l12_dequant(const Layer12Args &a, cudaStream_t s);
cudaError_t launch_l12_synth(const DeviceTables &t, const Synth12Args &a, cudaStream_t s);

}  // namespace mp3b
#endif
