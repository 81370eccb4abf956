/* examples/decode_file.c -- a minimp3 caller in plain C, unchanged except for the header it includes.
 *
 *   cc -std=c99 -Iinclude examples/decode_file.c -Lminimp3_b200 -lminimp3_b200 -Wl,-rpath,$PWD/minimp3_b200 -o decode_file
 *   ./decode_file in.mp3 [out.pcm]
 *
 * Decodes the file twice -- once with mp3dec_load_buf (one GPU batch) and once with the README frame loop over
 * mp3dec_decode_frame (minimp3 README "Usage") -- checks that both give the same number of samples for a
 * tag-free stream, and prints what the reference's test tool prints (minimp3_test.c:422): rate, samples.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "minimp3_b200.h"

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s in.mp3 [out.pcm]\n", argv[0]); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    fseek(f, 0, SEEK_END);
    long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *buf = (uint8_t *)malloc(size > 0 ? (size_t)size : 1);
    if (!buf || fread(buf, 1, (size_t)size, f) != (size_t)size) { fprintf(stderr, "read error\n"); return 2; }
    fclose(f);

    static mp3dec_t mp3d;
    mp3dec_file_info_t info;
    int ret = mp3dec_load_buf(&mp3d, buf, (size_t)size, &info, NULL, NULL);
    if (ret) { fprintf(stderr, "mp3dec_load_buf: %d\n", ret); return 1; }

    /* the frame loop of the reference's README */
    mp3dec_frame_info_t fi;
    static mp3d_sample_t pcm[MINIMP3_MAX_SAMPLES_PER_FRAME];
    size_t loop_samples = 0;
    const uint8_t *p = buf;
    long left = size;
    mp3dec_init(&mp3d);
    for (;;)
    {
        int samples = mp3dec_decode_frame(&mp3d, p, (int)left, pcm, &fi);
        if (!samples && !fi.frame_bytes) break;
        p += fi.frame_bytes;
        left -= fi.frame_bytes;
        loop_samples += (size_t)samples*fi.channels;
    }
    unsigned long long sum = 0;
    for (size_t i = 0; i < info.samples; i++) sum = sum*1099511628211ull + (unsigned short)info.buffer[i];
    printf("rate=%d channels=%d layer=%d samples=%zu frame_loop_samples=%zu kbps=%d fnv=%016llx\n", info.hz, info.channels,
           info.layer, info.samples, loop_samples, info.avg_bitrate_kbps, sum);
    if (argc > 2)
    {
        FILE *o = fopen(argv[2], "wb");
        if (!o) { perror(argv[2]); return 2; }
        fwrite(info.buffer, sizeof(mp3d_sample_t), info.samples, o);
        fclose(o);
    }
    free(info.buffer);
    free(buf);
    return 0;
}
