/*
 * mp3_hdr.cuh -- MPEG audio frame-header arithmetic, host and device (the host parser and the GPU frame scan,
 * k_scan.cu, must agree bit for bit).  Behaviour of the header macros / helpers of the reference
 * (/root/reference/minimp3.h:61-78, 264-314): validity, compatibility of two headers, bit rate, sample rate,
 * samples per frame, frame length and padding.
 */
#ifndef MP3B_MP3_HDR_CUH
#define MP3B_MP3_HDR_CUH
#include <stdint.h>

#if defined(__CUDACC__)
#define MP3H_HD __host__ __device__ __forceinline__
#else
#define MP3H_HD inline
#endif

namespace mp3b {

MP3H_HD bool hd_is_mono(const uint8_t *h) { return (h[3] & 0xC0) == 0xC0; }
MP3H_HD bool hd_is_crc(const uint8_t *h) { return !(h[1] & 1); }
MP3H_HD bool hd_is_mpeg1(const uint8_t *h) { return (h[1] & 0x8) != 0; }
MP3H_HD int hd_layer(const uint8_t *h) { return 4 - ((h[1] >> 1) & 3); }
MP3H_HD bool hd_is_free_format(const uint8_t *h) { return (h[2] & 0xF0) == 0; }

MP3H_HD bool hd_valid(const uint8_t *h)
{
    if (h[0] != 0xff) return false;
    const bool sync = (h[1] & 0xF0) == 0xF0 || (h[1] & 0xFE) == 0xE2;   /* MPEG-1/2, or MPEG-2.5 Layer III */
    const int layer_bits = (h[1] >> 1) & 3, bitrate = h[2] >> 4, sr = (h[2] >> 2) & 3;
    return sync && layer_bits != 0 && bitrate != 15 && sr != 3;
}

/* same version / layer, same sample rate, same free-format-ness (the protection bit may differ) */
MP3H_HD bool hd_compare(const uint8_t *h1, const uint8_t *h2)
{
    if (!hd_valid(h2)) return false;
    if ((h1[1] ^ h2[1]) & 0xFE) return false;
    if ((h1[2] ^ h2[2]) & 0x0C) return false;
    return hd_is_free_format(h1) == hd_is_free_format(h2);
}

/* ISO 11172-3 / 13818-3 2.4.2.3 bitrate index tables in kbit/s, halved to fit a byte; row = 3 x (MPEG-1) + layer
 * bits - 1 (layer bits: 1 = III, 2 = II, 3 = I).  One copy per address space: a table inside the function would be
 * rebuilt on the stack by every device call. */
#define MP3H_HALF_KBPS { \
        { 0, 4, 8, 12, 16, 20, 24, 28, 32, 40, 48, 56, 64, 72, 80 },              /* LSF layer III */ \
        { 0, 4, 8, 12, 16, 20, 24, 28, 32, 40, 48, 56, 64, 72, 80 },              /* LSF layer II  */ \
        { 0, 16, 24, 28, 32, 40, 48, 56, 64, 72, 80, 88, 96, 112, 128 },          /* LSF layer I   */ \
        { 0, 16, 20, 24, 28, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160 },         /* MPEG-1 layer III */ \
        { 0, 16, 24, 28, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192 },        /* MPEG-1 layer II  */ \
        { 0, 16, 32, 48, 64, 80, 96, 112, 128, 144, 160, 176, 192, 208, 224 } }   /* MPEG-1 layer I   */
static const uint8_t k_half_kbps_host[6][15] = MP3H_HALF_KBPS;
#if defined(__CUDACC__)
static __device__ const uint8_t k_half_kbps_dev[6][15] = MP3H_HALF_KBPS;
#endif

MP3H_HD unsigned hd_bitrate_kbps(const uint8_t *h)
{
    const int row = (hd_is_mpeg1(h) ? 3 : 0) + ((h[1] >> 1) & 3) - 1;
#if defined(__CUDA_ARCH__)
    return 2u*k_half_kbps_dev[row][h[2] >> 4];
#else
    return 2u*k_half_kbps_host[row][h[2] >> 4];
#endif
}

MP3H_HD unsigned hd_sample_rate_hz(const uint8_t *h)
{
    const int idx = (h[2] >> 2) & 3;
    unsigned hz = idx == 0 ? 44100u : idx == 1 ? 48000u : 32000u;
    if (!hd_is_mpeg1(h)) hz >>= 1;           /* MPEG-2 */
    if (!(h[1] & 0x10)) hz >>= 1;            /* MPEG-2.5 */
    return hz;
}

MP3H_HD unsigned hd_frame_samples(const uint8_t *h)
{
    if ((h[1] & 6) == 6) return 384;                  /* layer I */
    return ((h[1] & 14) == 2) ? 576 : 1152;           /* LSF layer III : everything else */
}

MP3H_HD int hd_frame_bytes(const uint8_t *h, int free_format_size)
{
    int bytes = (int)(hd_frame_samples(h)*hd_bitrate_kbps(h)*125/hd_sample_rate_hz(h));
    if ((h[1] & 6) == 6) bytes &= ~3;                 /* layer I slots are 4 bytes */
    return bytes ? bytes : free_format_size;
}

MP3H_HD int hd_padding(const uint8_t *h)
{
    if (!(h[2] & 0x2)) return 0;
    return ((h[1] & 6) == 6) ? 4 : 1;
}

}  // namespace mp3b
#endif
