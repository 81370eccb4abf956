/*
 * minimp3_b200.h -- C ABI of libminimp3_b200.so: a CUDA (sm_100a) Layer III decoder behind the
 * public API of lieff/minimp3, plus one batch entry point.
 *
 * Every declaration cites the reference interface it replaces (paths relative to the reference
 * checkout).  Types keep the reference's layout so existing callers and language bindings can switch
 * by relinking.  There is NO CPU decode path in this library: every PCM sample is produced by the
 * CUDA kernels; all entry points fail (MP3D_E_CUDA) when no usable device is present.
 */
#ifndef MINIMP3_B200_H
#define MINIMP3_B200_H
#include <stddef.h>
#include <stdint.h>

#define MINIMP3_MAX_SAMPLES_PER_FRAME (1152*2)          /* minimp3.h:11 */

/* minimp3_ex.h:32-36 */
#define MP3D_E_PARAM   -1
#define MP3D_E_MEMORY  -2
#define MP3D_E_IOERROR -3
#define MP3D_E_USER    -4
#define MP3D_E_DECODE  -5
/* new codes of this library */
#define MP3D_E_UNSUPPORTED -6   /* reserved: stream needs a decode stage that is not on the CUDA path */
#define MP3D_E_CUDA        -7   /* CUDA runtime error / no device */

/* minimp3.h:13-16 */
typedef struct
{
    int frame_bytes, frame_offset, channels, hz, layer, bitrate_kbps;
} mp3dec_frame_info_t;

/* minimp3.h:18-23 -- same size and field layout (6668 bytes).  mdct_overlap holds the same values as
 * the reference's; qmf_state holds the 15 carried synthesis slots in this library's own order. */
typedef struct
{
    float mdct_overlap[2][9*32], qmf_state[15*2*32];
    int reserv, free_format_bytes;
    unsigned char header[4], reserv_buf[511];
} mp3dec_t;

#ifdef MINIMP3_FLOAT_OUTPUT                               /* minimp3.h:30-35 */
typedef float mp3d_sample_t;
#define mp3dec_decode_frame mp3dec_decode_frame_f32
#define mp3dec_load_buf mp3dec_load_buf_f32
#else
typedef int16_t mp3d_sample_t;
#endif

/* minimp3_ex.h:38-43.  With float output selected at run time (batch opts / *_f32 entry points) buffer
 * really points at float samples. */
typedef struct
{
    mp3d_sample_t *buffer;
    size_t samples;        /* channels included, byte size = samples*sizeof(sample) */
    int channels, hz, layer, avg_bitrate_kbps;
} mp3dec_file_info_t;

/* minimp3_ex.h:91 */
typedef int (*MP3D_PROGRESS_CB)(void *user_data, size_t file_size, uint64_t offset, mp3dec_frame_info_t *info);

#ifdef __cplusplus
extern "C" {
#endif

/* ---- frame API: minimp3.h:29-36 ------------------------------------------------------------- */
void mp3dec_init(mp3dec_t *dec);
/* returns samples per channel (0 / 384 / 576 / 1152); info->frame_bytes = bytes to drop.  pcm may be NULL
 * (parse only).  Layers I, II and III are all decoded on the GPU.  A freshly initialised decoder that is handed a
 * buffer with many frames decodes them as one look-ahead window; calls that continue where the previous frame ended
 * are answered from it (MP3B_FRAME_LOOKAHEAD=0 turns this off). */
int mp3dec_decode_frame(mp3dec_t *dec, const uint8_t *mp3, int mp3_bytes, mp3d_sample_t *pcm, mp3dec_frame_info_t *info);
#ifndef MINIMP3_FLOAT_OUTPUT
int mp3dec_decode_frame_f32(mp3dec_t *dec, const uint8_t *mp3, int mp3_bytes, float *pcm, mp3dec_frame_info_t *info);
#endif
void mp3dec_f32_to_s16(const float *in, int16_t *out, int num_samples);   /* minimp3.h:34, 1808-1864 (runs on the GPU) */

/* ---- whole-buffer API: minimp3_ex.h:98,101 ----------------------------------------------------- */
int mp3dec_detect_buf(const uint8_t *buf, size_t buf_size);
/* info->buffer is malloc()ed, caller free()s (README.md:258-259).  Without progress_cb the stream is decoded as
 * one GPU batch; with progress_cb (minimp3_ex.h:52, 504-509) the frames come out of look-ahead windows so that the
 * callback sees every decoded frame in order and a non-zero return stops the load and is returned. */
int mp3dec_load_buf(mp3dec_t *dec, const uint8_t *buf, size_t buf_size, mp3dec_file_info_t *info, MP3D_PROGRESS_CB progress_cb, void *user_data);
#ifndef MINIMP3_FLOAT_OUTPUT
int mp3dec_load_buf_f32(mp3dec_t *dec, const uint8_t *buf, size_t buf_size, mp3dec_file_info_t *info, MP3D_PROGRESS_CB progress_cb, void *user_data);
#endif

#ifdef __cplusplus
}
#endif

/* ---- iterate / streaming / seek API: minimp3_ex.h:12-29, 45-124 ----------------------------------- */
#define MP3D_SEEK_TO_BYTE   0
#define MP3D_SEEK_TO_SAMPLE 1
#define MP3D_DO_NOT_SCAN    2
#define MP3D_ALLOW_MONO_STEREO_TRANSITION 4      /* always available here (run-time flag) */
#define MP3D_FLAGS_MASK 7
#define MINIMP3_PREDECODE_FRAMES 2
#define MINIMP3_IO_SIZE (128*1024)
#define MINIMP3_BUF_SIZE (16*1024)

typedef struct { const uint8_t *buffer; size_t size; } mp3dec_map_info_t;
typedef struct { uint64_t sample; uint64_t offset; } mp3dec_frame_t;
typedef struct { mp3dec_frame_t *frames; size_t num_frames, capacity; } mp3dec_index_t;
typedef size_t (*MP3D_READ_CB)(void *buf, size_t size, void *user_data);
typedef int (*MP3D_SEEK_CB)(uint64_t position, void *user_data);
typedef struct { MP3D_READ_CB read; void *read_data; MP3D_SEEK_CB seek; void *seek_data; } mp3dec_io_t;

/* minimp3_ex.h:74-88, same field order */
typedef struct
{
    mp3dec_t mp3d;
    mp3dec_map_info_t file;
    mp3dec_io_t *io;
    mp3dec_index_t index;
    uint64_t offset, samples, detected_samples, cur_sample, start_offset, end_offset;
    mp3dec_frame_info_t info;
    mp3d_sample_t buffer[MINIMP3_MAX_SAMPLES_PER_FRAME];
    size_t input_consumed, input_filled;
    int is_file, flags, vbr_tag_found, indexes_built;
    int free_format_bytes;
    int buffer_samples, buffer_consumed, to_skip, start_delay;
    int last_error;
} mp3dec_ex_t;

typedef int (*MP3D_ITERATE_CB)(void *user_data, const uint8_t *frame, int frame_size, int free_format_bytes,
                               size_t buf_size, uint64_t offset, mp3dec_frame_info_t *info);

#ifdef MINIMP3_FLOAT_OUTPUT
#define mp3dec_ex_open_buf mp3dec_ex_open_buf_f32
#define mp3dec_ex_open_cb mp3dec_ex_open_cb_f32
#define mp3dec_ex_open mp3dec_ex_open_f32
#define mp3dec_ex_close mp3dec_ex_close_f32
#define mp3dec_ex_seek mp3dec_ex_seek_f32
#define mp3dec_ex_read_frame mp3dec_ex_read_frame_f32
#define mp3dec_ex_read mp3dec_ex_read_f32
#define mp3dec_load_cb mp3dec_load_cb_f32
#define mp3dec_load mp3dec_load_f32
#endif

#ifdef __cplusplus
extern "C" {
#endif
int mp3dec_detect_cb(mp3dec_io_t *io, uint8_t *buf, size_t buf_size);                                  /* minimp3_ex.h:99 */
int mp3dec_load_cb(mp3dec_t *dec, mp3dec_io_t *io, uint8_t *buf, size_t buf_size, mp3dec_file_info_t *info,
                   MP3D_PROGRESS_CB progress_cb, void *user_data);                                     /* :102 */
int mp3dec_iterate_buf(const uint8_t *buf, size_t buf_size, MP3D_ITERATE_CB callback, void *user_data); /* :104 */
int mp3dec_iterate_cb(mp3dec_io_t *io, uint8_t *buf, size_t buf_size, MP3D_ITERATE_CB callback, void *user_data); /* :105 */
int mp3dec_ex_open_buf(mp3dec_ex_t *dec, const uint8_t *buf, size_t buf_size, int flags);              /* :107 */
int mp3dec_ex_open_cb(mp3dec_ex_t *dec, mp3dec_io_t *io, int flags);                                   /* :108 */
void mp3dec_ex_close(mp3dec_ex_t *dec);                                                                /* :109 */
int mp3dec_ex_seek(mp3dec_ex_t *dec, uint64_t position);                                               /* :110 */
size_t mp3dec_ex_read_frame(mp3dec_ex_t *dec, mp3d_sample_t **buf, mp3dec_frame_info_t *frame_info, size_t max_samples); /* :111 */
size_t mp3dec_ex_read(mp3dec_ex_t *dec, mp3d_sample_t *buf, size_t samples);                           /* :112 */
/* stdio flavours (:115-118): the file is read into memory and handed to the buffer flavours */
int mp3dec_detect(const char *file_name);
int mp3dec_load(mp3dec_t *dec, const char *file_name, mp3dec_file_info_t *info, MP3D_PROGRESS_CB progress_cb, void *user_data);
int mp3dec_iterate(const char *file_name, MP3D_ITERATE_CB callback, void *user_data);
int mp3dec_ex_open(mp3dec_ex_t *dec, const char *file_name, int flags);
#ifdef __cplusplus
}
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- batch API (new) --------------------------------------------------------------- */
typedef struct
{
    const uint8_t *buf;
    size_t size;
} mp3dec_batch_stream_t;

#define MP3D_BATCH_OUT_MALLOC 0   /* infos[i].buffer = malloc()ed host copy, caller free()s (drop-in) */
#define MP3D_BATCH_OUT_PINNED 1   /* infos[i].buffer points into a pinned slab owned by the library */
#define MP3D_BATCH_OUT_DEVICE 2   /* infos[i].buffer is a device pointer into the library's PCM buffer */
/* Lifetime of PINNED / DEVICE outputs of mp3dec_decode_batch_cuda: until the caller's next PINNED / DEVICE call on the
 * same device, or mp3dec_batch_release_outputs(device).  Nothing else the library does (malloc-output batch calls,
 * mp3dec_load_buf, the frame and streaming APIs, other threads' calls in malloc mode) touches them.  Outputs of a
 * plan (mp3dec_batch_plan_fetch) live until mp3dec_batch_plan_destroy. */

typedef struct
{
    int device;                       /* CUDA device ordinal */
    int float_output;                 /* 0: int16 PCM, 1: float PCM (MINIMP3_FLOAT_OUTPUT semantics) */
    int raw_frames;                   /* 1: plain mp3dec_decode_frame loop, no tag skipping / VBR-tag trimming */
    int allow_mono_stereo_transition; /* MINIMP3_ALLOW_MONO_STEREO_TRANSITION */
    int host_threads;                 /* 0: all online cores */
    int output;                       /* MP3D_BATCH_OUT_* */
    void *cuda_stream;                /* cudaStream_t to run on, NULL: library-owned stream */
    int n_devices;                    /* 0 / 1: `device` alone.  2..8: mp3dec_decode_batch_cuda shards the streams over */
    int devices[8];                   /* devices[0 .. n_devices-1] (greedy LPT on input bytes, no collective); */
                                      /* `device` is ignored then.  Plans (mp3dec_batch_plan_*) are single-device. */
    int gpu_frame_scan;               /* frame sync + reservoir accounting on the GPU (mp3dec_decode_batch_cuda only): */
                                      /* 0 auto (on when pinned_inputs is given: without it the host's staging copy */
                                      /* dominates and the scan only adds latency), 1 always, -1 never.  Results do not depend on it. */
    const void *pinned_inputs;        /* optional: a page-locked (cudaHostAlloc / cudaHostRegister) region that holds the */
    size_t pinned_inputs_bytes;       /* stream buffers.  With the GPU frame scan the streams are then uploaded straight */
                                      /* from it (no staging copy); NULL / 0: any memory, staged through the library's slab. */
} mp3dec_batch_opts_t;

/* Decodes n_streams independent buffers; per-stream result and return code (infos_out[i], rets_out[i],
 * may be NULL) are those of mp3dec_load_buf(&dec, buf, size, &info, 0, 0).  Returns 0, or MP3D_E_PARAM /
 * MP3D_E_MEMORY / MP3D_E_CUDA if the batch as a whole could not run. */
int mp3dec_decode_batch_cuda(const mp3dec_batch_stream_t *streams, int n_streams, mp3dec_file_info_t *infos_out,
                             int *rets_out, const mp3dec_batch_opts_t *opts);

void mp3dec_batch_release_outputs(int device);

/* Where the time of the calling thread's last mp3dec_decode_batch_cuda() went (measurement aid). */
typedef struct
{
    double total_ms;             /* whole call */
    double host_parse_ms;        /* host stage (frame walk + staging copies), summed over the chunks of the call */
    double issue_ms;             /* until the last chunk's kernels and read-back were queued */
    double finish_ms;            /* waiting for the read-backs + output hand-over (malloc copies) */
    uint64_t h2d_bytes, d2h_bytes;
    int n_chunks, n_devices;
    int scanned_streams;         /* streams whose frames were found by the GPU frame scan (the rest: host walker) */
} mp3dec_batch_call_stats_t;
void mp3dec_batch_last_call_stats(mp3dec_batch_call_stats_t *out);

/* Split form of the batch call, for callers that keep inputs resident (and for measurement):
 * plan = host parse + H2D; run = kernels only (asynchronous on the stream); fetch = D2H + infos. */
typedef struct mp3dec_batch_plan mp3dec_batch_plan_t;

typedef struct
{
    uint64_t input_bytes;        /* bytes of frame payload staged to the device */
    uint64_t samples;            /* channel-inclusive samples the kernels write */
    uint64_t n_granule_channels, n_tiles, n_istereo_granules, n_frames;
    uint64_t h2d_bytes;          /* bytes copied host to device by plan creation */
    int kernels_per_run;         /* kernel launches one mp3dec_batch_plan_run() issues */
    double host_parse_ms, h2d_ms;
    int kernels_prep;            /* own kernels of the plan's GPU preparation (side-info parse, lane-order keys) */
} mp3dec_batch_stats_t;

mp3dec_batch_plan_t *mp3dec_batch_plan_create(const mp3dec_batch_stream_t *streams, int n_streams,
                                              const mp3dec_batch_opts_t *opts, int *err);
int mp3dec_batch_plan_run(mp3dec_batch_plan_t *plan);
/* measurement aid: one stage only (0 entropy, 1 intensity stereo, 2 transform, 3 Layer I/II, 4 the plan's GPU
 * preparation: side-info parse + lane order, which plan creation already ran once) */
int mp3dec_batch_plan_run_stage(mp3dec_batch_plan_t *plan, int stage);
int mp3dec_batch_plan_sync(mp3dec_batch_plan_t *plan);
int mp3dec_batch_plan_fetch(mp3dec_batch_plan_t *plan, mp3dec_file_info_t *infos_out, int *rets_out);
void mp3dec_batch_plan_stats(const mp3dec_batch_plan_t *plan, mp3dec_batch_stats_t *out);
void *mp3dec_batch_plan_device_pcm(mp3dec_batch_plan_t *plan);
void mp3dec_batch_plan_destroy(mp3dec_batch_plan_t *plan);

/* Test taps (parity tests): device-side intermediates of the last mp3dec_batch_plan_run(), copied to host.
 * which: 0 quantised lines int16[n_gc][576], 1 scalefactors float[n_gc][40], 2 requantised spectra
 * float[n_gc][576] (after MS, zero beyond nzl), 3 subband samples float[n_gc][576], 4 GcDesc records,
 * 5 ist_pos uint8[n_gc][40], 8 the entropy kernel's lane order uint32[n_order] (readable without enabling taps).
 * Taps must be enabled before run. */
int mp3dec_batch_plan_enable_taps(mp3dec_batch_plan_t *plan);
int mp3dec_batch_plan_read_tap(mp3dec_batch_plan_t *plan, int which, void *dst, size_t dst_bytes);
/* Test tap: |x|^(4/3) for x = 0 .. n-1 (n <= 8207) exactly as the entropy kernel's requantiser evaluates it
 * (L3_pow_43, minimp3.h:716-740), computed on the current device.  0 or MP3D_E_MEMORY / MP3D_E_PARAM. */
int mp3dec_b200_pow43_device(float *dst, int n);

/* Multi-GPU sharding helper (what opts.devices uses inside the batch call): assigns streams to n_shards by
 * greedy longest-processing-time on input bytes; shard_of[i] receives the shard of stream i.  No data-path
 * collective exists: shards are decoded independently. */
void mp3dec_batch_shard(const size_t *sizes, int n_streams, int n_shards, int *shard_of);

const char *mp3dec_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MINIMP3_B200_H */
