/* oracle/mp3_oracle.h -- TEST INFRASTRUCTURE ONLY (see mp3_oracle.c). */
#ifndef MP3_ORACLE_H
#define MP3_ORACLE_H
#include <stdint.h>

typedef struct
{
    long cap;        /* capacity in granule-channels */
    int16_t *q;      /* [cap][576] Huffman-decoded signed integers            (or NULL) */
    float *scf;      /* [cap][40]  scalefactors as the requantiser uses them  (or NULL) */
    float *xr;       /* [cap][576] requantised lines before stereo            (or NULL) */
    float *sub;      /* [cap][576] subband samples after IMDCT + inversion    (or NULL) */
} orc_taps;

#ifdef __cplusplus
extern "C" {
#endif
/* Decodes every frame of a stream the way a mp3dec_init + mp3dec_decode_frame loop does.
 * Returns channel-inclusive samples.  pcm16 / pcmf may be NULL. */
long orc_decode_stream(const uint8_t *buf, long size, int16_t *pcm16, float *pcmf, long pcm_cap, orc_taps *taps,
                       long *n_gc_out);
#ifdef __cplusplus
}
#endif
#endif
