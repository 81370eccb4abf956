/*
 * tests/fuzz_host_parse.cpp -- TEST-ONLY sanitizer harness for the product's host parser (csrc/host_parse.cpp).
 *
 * The reference fuzzes its decoder through libFuzzer (fuzzing/fuzz.c: mp3dec_decode_frame over the input) and ships
 * an AFL recipe (scripts/fuzz.sh).  The host parser is where this library touches untrusted bytes before anything
 * reaches the GPU, and the only place a sanitizer can run on this pool (compute-sanitizer is closed), so it gets the
 * same treatment:
 *
 *   LLVMFuzzerTestOneInput()   the libFuzzer entry (clang -fsanitize=fuzzer,address,undefined ... this file
 *                              csrc/host_parse.cpp), usable as is where clang is available;
 *   main()                     -DMP3B_FUZZ_STANDALONE: a self-contained mutation driver for toolchains without
 *                              libFuzzer (g++ -fsanitize=address,undefined): seeds = the files given on the command
 *                              line, deterministic xorshift mutations (byte damage, truncation, splices, deletions,
 *                              duplicated ranges, planted sync words and tags), runs for --seconds N.
 *
 * One input exercises: tag skipping + first-frame probe, the whole-stream walk in load_buf and raw-frame mode with
 * host and device side-info records (exact-size heap buffers: the walker promises to stay inside size + 16 bytes of
 * main data), the single-frame parser loop as the frame API drives it, detect_buf, the side-info -> descriptor
 * expansion the GPU kernel runs (frame_descs, host build), and the invariants the device code relies on (bit offsets
 * inside the main-data buffer, tiles covering the granules, lane order a permutation).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <string>
#include <vector>

#include "host_parse.h"
#include "l3_sideinfo.cuh"

using namespace mp3b;

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "invariant failed: %s (%s:%d)\n", #c, __FILE__, __LINE__); abort(); } } while (0)

static void walk_and_check(const uint8_t *data, size_t size, bool raw, bool allow, bool device_side_info)
{
    StreamPlan plan;
    WalkOptions opt;
    opt.raw_frames = raw;
    opt.allow_mono_stereo = allow;
    opt.device_side_info = device_side_info;
    /* exactly what the batch call gives the walker: size + 16 bytes, on the heap so that ASan sees any overrun */
    uint8_t *md = (uint8_t *)malloc(size + 16);
    CHECK(md);
    uint8_t *in = (uint8_t *)malloc(size ? size : 1);      /* exact-size copy: over-reads of the input are caught too */
    CHECK(in);
    if (size) memcpy(in, data, size);
    walk_stream(in, size, md, 0, opt, &plan);
    CHECK(plan.main_data_bytes <= size);
    CHECK(plan.out_samples <= plan.samples_decoded);
    CHECK(plan.out_skip <= plan.samples_decoded);
    const uint64_t md_bits = (uint64_t)plan.main_data_bytes*8u;
    uint64_t tile_samples = 0;
    for (const TileDesc &t : plan.tiles)
    {
        CHECK(t.n_gr >= 1 && t.n_gr <= (t.nch == 1 ? MP3B_TILE_GRANULES_MONO : MP3B_TILE_GRANULES) && (t.nch == 1 || t.nch == 2));
        CHECK((uint64_t)t.gc_first + (uint64_t)t.n_gr*t.nch <= plan.n_gc);
        for (int c = 0; c < 2; c++)
            for (int h = 0; h < 2; h++) CHECK(t.halo[c][h] < (int64_t)plan.n_gc);
        CHECK(t.pcm_off == tile_samples || !plan.tiles12.empty());
        tile_samples = t.pcm_off + 576ull*t.n_gr*t.nch;
    }
    if (device_side_info)
    {
        CHECK(plan.gc_key.size() == plan.n_gc);
        for (const FrameRec &r : plan.frame_recs)
        {
            CHECK(r.md_bit <= md_bits);
            CHECK((uint64_t)r.gc_first + (uint64_t)r.n_gr*r.nch <= plan.n_gc);
            GcDesc d[4];
            const int n = frame_descs(r, d);           /* the expansion k_l3_side_info runs */
            CHECK(n == r.n_gr*r.nch);
            for (int k = 0; k < n; k++)
            {
                CHECK(d[k].bit_off <= md_bits && d[k].part_23_length <= 4095 && d[k].big_values <= 288);
                CHECK(d[k].n_long_sfb + d[k].n_short_sfb <= 39 && d[k].sr_idx < 8);
            }
        }
    } else
    {
        CHECK(plan.gcs.size() == plan.n_gc);
        for (const GcDesc &d : plan.gcs) CHECK(d.bit_off <= md_bits && d.big_values <= 288 && d.sr_idx < 8);
    }
    std::vector<uint8_t> seen(plan.n_gc, 0);
    for (uint32_t k : plan.order)
    {
        CHECK(k < plan.n_gc && !seen[k]);
        seen[k] = 1;
    }
    for (const L12FrameDesc &f : plan.l12_frames) CHECK(f.bit_off <= md_bits && f.n_blocks >= 1 && f.n_blocks <= 3);
    free(in);
    free(md);
}

static void frame_loop(const uint8_t *data, size_t size)
{
    /* the README loop as the frame API parses it: bounded view lengths on top, as a streaming caller would pass */
    uint8_t *in = (uint8_t *)malloc(size ? size : 1);
    CHECK(in);
    if (size) memcpy(in, data, size);
    SyncState st;
    FrameParse fp;
    memset(&fp.info, 0, sizeof(fp.info));
    size_t pos = 0;
    for (int guard = 0; guard < 100000; guard++)
    {
        const size_t left = size - pos;
        const int view = (int)std::min<size_t>(left, (guard & 7) == 7 ? 3000 : left);
        parse_frame(&st, in + pos, view, &fp, (guard & 15) == 3, (guard & 1) != 0);
        CHECK(fp.info.frame_bytes >= 0 && (size_t)fp.info.frame_bytes <= (size_t)view);
        if (fp.payload) CHECK(fp.payload >= in + pos && fp.payload + fp.payload_bytes <= in + pos + view);
        CHECK(st.reserv >= 0 && st.reserv <= 511);
        if (!fp.info.frame_bytes) break;
        pos += (size_t)fp.info.frame_bytes;
    }
    free(in);
}

extern "C" int LLVMFuzzerTestOneInput(const uint8_t *data, size_t size)
{
    if (size > (1u << 20)) size = 1u << 20;
    walk_and_check(data, size, false, false, true);
    walk_and_check(data, size, false, true, false);
    walk_and_check(data, size, true, false, true);
    walk_and_check(data, size, true, true, false);
    frame_loop(data, size);
    {
        uint8_t *in = (uint8_t *)malloc(size ? size : 1);
        CHECK(in);
        if (size) memcpy(in, data, size);
        const int r = detect_buf(in, size);
        CHECK(r == 0 || r == -4 || r == -1);
        const uint8_t *p = in;
        size_t n = size;
        StreamHead head;
        if (open_stream_head(&p, &n, &head)) CHECK(p >= in && p + n <= in + size && head.first.frame_bytes > 0);
        free(in);
    }
    return 0;
}

#ifdef MP3B_FUZZ_STANDALONE
static uint64_t g_rng = 0x9E3779B97F4A7C15ull;
static uint64_t rnd()
{
    g_rng ^= g_rng << 13; g_rng ^= g_rng >> 7; g_rng ^= g_rng << 17;
    return g_rng;
}
static size_t rnd_below(size_t n) { return n ? (size_t)(rnd() % n) : 0; }

static void mutate(std::vector<uint8_t> &b, const std::vector<std::vector<uint8_t>> &seeds)
{
    static const uint8_t sync[][2] = { {0xFF, 0xFB}, {0xFF, 0xFA}, {0xFF, 0xF3}, {0xFF, 0xE3}, {0xFF, 0xFD}, {0xFF, 0xFC}, {0xFF, 0xF2} };
    const int rounds = 1 + (int)rnd_below(6);
    for (int r = 0; r < rounds; r++)
    {
        switch (rnd_below(9))
        {
        case 0: for (int k = 0, n = 1 + (int)rnd_below(16); k < n && !b.empty(); k++) b[rnd_below(b.size())] = (uint8_t)rnd(); break;
        case 1: if (!b.empty()) b.resize(rnd_below(b.size() + 1)); break;
        case 2: if (b.size() > 8) { size_t p = rnd_below(b.size() - 4), n = 1 + rnd_below(std::min<size_t>(600, b.size() - p)); b.erase(b.begin() + p, b.begin() + p + n); } break;
        case 3: if (b.size() > 8) { size_t p = rnd_below(b.size() - 4), n = 1 + rnd_below(std::min<size_t>(600, b.size() - p)); std::vector<uint8_t> c(b.begin() + p, b.begin() + p + n); b.insert(b.begin() + p, c.begin(), c.end()); } break;
        case 4: { const std::vector<uint8_t> &o = seeds[rnd_below(seeds.size())]; if (!o.empty()) { size_t p = rnd_below(o.size()), n = rnd_below(std::min<size_t>(4000, o.size() - p)); b.insert(b.begin() + rnd_below(b.size() + 1), o.begin() + p, o.begin() + p + n); } } break;
        case 5: for (int k = 0, n = 1 + (int)rnd_below(12); k < n && b.size() > 4; k++) { size_t p = rnd_below(b.size() - 1); const uint8_t *s = sync[rnd_below(7)]; b[p] = s[0]; b[p + 1] = s[1]; } break;
        case 6: if (b.size() > 1) b.erase(b.begin(), b.begin() + rnd_below(std::min<size_t>(3000, b.size()))); break;
        case 7: { static const char *tags[] = { "ID3\x03\x00\x00\x00\x00\x0f\x7f", "TAG", "APETAGEX", "Xing\x00\x00\x00\x0f", "Info\x00\x00\x00\x01", "TAG+" };
                  const char *t = tags[rnd_below(6)]; size_t n = strlen(t) ? strlen(t) : 4; size_t p = rnd_below(3) == 0 ? 0 : rnd_below(b.size() + 1);
                  if (rnd_below(2) && b.size() >= 128) p = b.size() - 128;
                  b.insert(b.begin() + std::min(p, b.size()), (const uint8_t *)t, (const uint8_t *)t + n); } break;
        default: if (b.size() > 4) { size_t p = rnd_below(b.size() - 4); b[p + 2] ^= (uint8_t)(1u << rnd_below(8)); b[p + 3] ^= (uint8_t)(1u << rnd_below(8)); } break;   /* header bit flips */
        }
    }
}

int main(int argc, char **argv)
{
    double seconds = 10.0;
    std::vector<std::vector<uint8_t>> seeds;
    for (int i = 1; i < argc; i++)
    {
        if (!strcmp(argv[i], "--seconds") && i + 1 < argc) { seconds = atof(argv[++i]); continue; }
        if (!strcmp(argv[i], "--seed") && i + 1 < argc) { g_rng ^= strtoull(argv[++i], 0, 10)*0x2545F4914F6CDD1Dull; continue; }
        FILE *f = fopen(argv[i], "rb");
        if (!f) { fprintf(stderr, "cannot open %s\n", argv[i]); return 2; }
        std::vector<uint8_t> b;
        uint8_t tmp[65536];
        size_t n;
        while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0) b.insert(b.end(), tmp, tmp + n);
        fclose(f);
        seeds.push_back(b);
    }
    if (seeds.empty()) seeds.push_back(std::vector<uint8_t>(4096, 0xff));
    for (const auto &s : seeds) LLVMFuzzerTestOneInput(s.data(), s.size());      /* the seeds themselves first */
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    unsigned long runs = 0;
    for (;;)
    {
        std::vector<uint8_t> b = seeds[rnd_below(seeds.size())];
        if (b.size() > 60000 && rnd_below(4)) b.resize(20000 + rnd_below(40000));        /* keep iterations short */
        mutate(b, seeds);
        LLVMFuzzerTestOneInput(b.data(), b.size());
        runs++;
        clock_gettime(CLOCK_MONOTONIC, &t1);
        if ((double)(t1.tv_sec - t0.tv_sec) + 1e-9*(double)(t1.tv_nsec - t0.tv_nsec) >= seconds) break;
    }
    printf("fuzz_host_parse: %lu mutated inputs + %zu seeds, no sanitizer report, no invariant failure\n", runs, seeds.size());
    return 0;
}
#endif
