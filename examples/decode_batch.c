/* examples/decode_batch.c -- many files, one call: mp3dec_decode_batch_cuda from plain C.
 *
 *   cc -std=c99 -Iinclude examples/decode_batch.c -Lminimp3_b200 -lminimp3_b200 -Wl,-rpath,$PWD/minimp3_b200 -o decode_batch
 *   ./decode_batch a.mp3 b.mp3 ...            (each file may be listed several times)
 *
 * Per file it prints what mp3dec_load_buf would have returned for it: return code, rate, channels, samples.
 */
#include <stdio.h>
#include <stdlib.h>

#include "minimp3_b200.h"

static uint8_t *slurp(const char *name, size_t *size)
{
    FILE *f = fopen(name, "rb");
    if (!f) { perror(name); exit(2); }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *p = (uint8_t *)malloc(n > 0 ? (size_t)n : 1);
    if (!p || fread(p, 1, (size_t)n, f) != (size_t)n) { fprintf(stderr, "%s: read error\n", name); exit(2); }
    fclose(f);
    *size = (size_t)n;
    return p;
}

int main(int argc, char **argv)
{
    int n = argc - 1, i;
    if (n < 1) { fprintf(stderr, "usage: %s file.mp3 ...\n", argv[0]); return 2; }
    mp3dec_batch_stream_t *in = (mp3dec_batch_stream_t *)calloc((size_t)n, sizeof(*in));
    mp3dec_file_info_t *out = (mp3dec_file_info_t *)calloc((size_t)n, sizeof(*out));
    int *rets = (int *)calloc((size_t)n, sizeof(int));
    for (i = 0; i < n; i++) in[i].buf = slurp(argv[i + 1], &in[i].size);

    mp3dec_batch_opts_t opt = { 0 };
    opt.output = MP3D_BATCH_OUT_PINNED;      /* results stay valid until the next batch call on this device */
    int err = mp3dec_decode_batch_cuda(in, n, out, rets, &opt);
    if (err) { fprintf(stderr, "mp3dec_decode_batch_cuda: %d\n", err); return 1; }
    for (i = 0; i < n; i++)
    {
        long long sum = 0;
        size_t k;
        for (k = 0; k < out[i].samples; k++) sum += out[i].buffer[k];
        printf("%s ret=%d rate=%d channels=%d samples=%zu sum=%lld\n", argv[i + 1], rets[i], out[i].hz, out[i].channels,
               out[i].samples, sum);
    }
    for (i = 0; i < n; i++) free((void *)in[i].buf);
    free(in); free(out); free(rets);
    return 0;
}
