/*
 * k_scan.cu -- frame sync, side-info validity, bit-reservoir accounting and tile planning on the GPU
 * (the rest of what the host walker does per frame).
 *
 * What stays on the host per stream: tag skipping and the first-frame search (open_stream_head / find_frame --
 * sequential pattern search over arbitrary bytes, once per stream) and one memcpy of the stream into pinned staging.
 * Everything per FRAME moves here, for streams that are "regular": Layer III, not free format, every frame found
 * exactly where the frame before it ends, same channel count throughout.  That is the reference's own fast path
 * (mp3dec_decode_frame with a matching next header, minimp3.h:1719-1728); the first irregularity -- junk between
 * frames, an incompatible or damaged header, side info the reference rejects, a channel-mode change, a short tail --
 * marks the stream for the host walker (csrc/host_parse.cpp), which reproduces the reference's resync behaviour
 * byte for byte.  One irregularity is ordinary -- a stream that ends in a cut frame or a few bytes of junk: the scan
 * stops there (SCAN_TAIL) and the host repeats the reference's search on that short tail alone; only if it finds a
 * frame in it does the stream go to the walker.  Nothing is decoded differently: both routes produce the same records.
 *
 *   k_stream_scan   one thread per stream walks its frames: header chain (hd_compare with the previous and the next
 *                   header), CRC skip, the reduced side-info read (main_data_begin, part_23 lengths, the reject
 *                   conditions: read_side_info_lite_t), reservoir fill exactly as L3_restore/save_reservoir
 *                   (minimp3.h:1212-1236).  Writes one ScanFrame per frame and a summary per stream.
 *   (host)          prefix sums over the summaries: where every stream's records, descriptors, tiles and PCM go
 *   k_stream_emit   one thread per ScanFrame, no sequential dependence left: FrameRec (raw side-info bytes for
 *                   k_l3_side_info), lane-order entries, intensity-stereo list
 *   k_payload_gather one warp per frame: the frame payloads of a stream concatenated (the reservoir's byte stream)
 *   k_stream_tiles  one thread per tile: closed form, the granules of a regular stream are evenly spaced
 */
#include <cuda_runtime.h>

#include "kernels.h"
#include "l3_sideinfo.cuh"
#include "mp3_hdr.cuh"

namespace mp3b {

namespace {

/* The first 40 bytes of a frame (header, CRC, side info) in registers: five 64-bit words in bitstream order (bit 0 =
 * most significant bit of byte 0), fetched with aligned 32-bit loads whatever the frame's alignment. */
struct FrameHead
{
    unsigned long long w[5];
    __device__ __forceinline__ void load(const uint32_t *__restrict__ raw_words, uint64_t byte_addr)
    {
        const uint32_t *src = raw_words + (byte_addr >> 2);
        const uint32_t sh = (uint32_t)(byte_addr & 3)*8u;
        uint32_t v[11];
#pragma unroll
        for (int k = 0; k < 11; k++) v[k] = __ldg(src + k);
#pragma unroll
        for (int k = 0; k < 5; k++)
        {
            const uint32_t lo = __funnelshift_r(v[2*k], v[2*k + 1], sh), hi = __funnelshift_r(v[2*k + 1], v[2*k + 2], sh);
            w[k] = (unsigned long long)__byte_perm(lo, 0, 0x0123) << 32 | __byte_perm(hi, 0, 0x0123);
        }
    }
    __device__ __forceinline__ void header(uint8_t *h) const
    {
        h[0] = (uint8_t)(w[0] >> 56); h[1] = (uint8_t)(w[0] >> 48); h[2] = (uint8_t)(w[0] >> 40); h[3] = (uint8_t)(w[0] >> 32);
    }
    __device__ __forceinline__ unsigned long long word(int i) const
    {
        return i == 0 ? w[0] : i == 1 ? w[1] : i == 2 ? w[2] : i == 3 ? w[3] : i == 4 ? w[4] : 0ull;
    }
    /* n <= 32 bits at bit position pos (pos + n <= 320) */
    __device__ __forceinline__ uint32_t bits(int pos, int n) const
    {
        const int i = pos >> 6, s = pos & 63;
        const unsigned long long hi = word(i), lo = word(i + 1);
        const unsigned long long v = s ? (hi << s) | (lo >> (64 - s)) : hi;
        return (uint32_t)(v >> (64 - n));
    }
};

/* the reader read_side_info_lite_t expects (get / pos / limit), over the bytes behind the header */
struct HeadBits
{
    const FrameHead *f;
    int pos, limit;
    __device__ __forceinline__ uint32_t get(int n)
    {
        const int s = pos;
        pos += n;
        if (pos > limit || n == 0) return 0;
        return f->bits(32 + s, n);
    }
};

__global__ void __launch_bounds__(64) k_stream_scan(ScanArgs a)
{
    const uint32_t s = blockIdx.x*blockDim.x + threadIdx.x;
    if (s >= a.n_streams) return;
    const ScanStream st = a.streams[s];
    const uint32_t *raw_words = reinterpret_cast<const uint32_t *>(a.raw);
    ScanFrame *out = a.frames + st.slot_base;
    ScanSummary sum;
    sum.n_frames = sum.n_dec = sum.n_ist = sum.kbps_sum = 0;
    sum.md_len = 0;
    sum.status = SCAN_OK;
    sum.pad = 0; sum.tail_pos = 0; sum.pad2 = 0;
    uint8_t prev[4] = { st.first_hdr[0], st.first_hdr[1], st.first_hdr[2], st.first_hdr[3] };
    const int nch0 = hd_is_mono(prev) ? 1 : 2;
    const int n_gr = hd_is_mpeg1(prev) ? 2 : 1;
    sum.nch = (uint8_t)nch0;
    sum.n_gr = (uint8_t)n_gr;
    int reserv = 0;
    uint32_t p = 0;
    FrameHead cur;
    if (st.size > 4) cur.load(raw_words, st.src_off);
    while (sum.status == SCAN_OK)
    {
        const uint32_t left = st.size - p;
        if (left <= 4) break;                              /* the reference finds nothing in so few bytes */
        uint8_t f[4];
        cur.header(f);
        if (!hd_compare(prev, f)) { sum.status = SCAN_IRREGULAR; break; }
        const int fs = hd_frame_bytes(f, 0) + hd_padding(f);
        /* the next frame's head: its header closes this frame's check, its bytes are the next round's input; the
         * frames after it are asked for now so that they come from L2 */
        FrameHead nxt;
        const bool last = (uint32_t)fs == left;
        if (!last)
        {
            /* the frame does not fit, or nothing can follow it: the reference searches the remaining bytes afresh
             * (and, as a rule, finds nothing); the host repeats exactly that search on the tail */
            if ((uint32_t)fs + 4 > left) { sum.status = SCAN_TAIL; break; }
            nxt.load(raw_words, st.src_off + p + (uint32_t)fs);
            if (p + 4u*(uint32_t)fs < st.size)
                asm volatile("prefetch.global.L2 [%0];" :: "l"(a.raw + st.src_off + p + 4u*(uint32_t)fs));
        }
        const int have = (int)min(left, 40u);
        const int nch = hd_is_mono(f) ? 1 : 2;
        if (nch != nch0) { sum.status = SCAN_IRREGULAR; break; }     /* mono/stereo switch: host decides what follows */
        const int crc = hd_is_crc(f) ? 2 : 0;
        const int si_bytes = hd_is_mpeg1(f) ? (nch == 1 ? 17 : 32) : (nch == 1 ? 9 : 17);
        if (4 + crc + si_bytes > fs || 4 + crc + si_bytes > have) { sum.status = SCAN_IRREGULAR; break; }
        HeadBits bs;
        bs.f = &cur;
        bs.limit = (fs - 4)*8;                             /* the frame, as the reference's reader has it */
        bs.pos = crc*8;
        GranuleSide side[4];
        const int mdb = read_side_info_lite_t(&bs, side, f);
        if (mdb < 0 || bs.pos > bs.limit) { sum.status = SCAN_IRREGULAR; break; }   /* reference re-initialises here */
        const int payload_off = 4 + crc + si_bytes;
        const int payload = fs - payload_off;
        /* L3_restore_reservoir / L3_save_reservoir, byte counts only */
        const bool ok = reserv >= mdb;
        int bits = 0;
        if (ok)
            for (int k = 0; k < n_gr*nch; k++) bits += side[k].part_23_length;
        const int remains = min(reserv, mdb) + payload - (bits + 7)/8;
        reserv = remains > 511 ? 511 : remains;
        if (!last)
        {
            /* what follows is not a frame of this stream: same search from here, if the rest is short (junk at the end
             * of a file); anything longer is the host walker's business */
            uint8_t nx[4];
            nxt.header(nx);
            if (!hd_compare(f, nx)) { sum.status = left <= SCAN_TAIL_MAX ? SCAN_TAIL : SCAN_IRREGULAR; break; }
        }
        if (sum.n_frames >= st.slot_cap) { sum.status = SCAN_IRREGULAR; break; }
        ScanFrame r;
        r.pos = p;
        r.frame_size = (uint16_t)fs;
        r.payload_off = (uint8_t)payload_off;
        r.flags = (uint8_t)((ok ? SCANF_DECODABLE : 0) | (((f[3] & 0x10) && nch == 2) ? SCANF_ISTEREO : 0));
        r.main_data_begin = (uint16_t)mdb;
        r.pad = 0;
        r.md_before = sum.md_len;
        r.dec_index = sum.n_dec;
        r.ist_before = sum.n_ist;
        out[sum.n_frames] = r;
        sum.n_frames++;
        sum.md_len += (uint32_t)payload;
        if (ok)
        {
            sum.n_dec++;
            sum.kbps_sum += hd_bitrate_kbps(f);
            if (r.flags & SCANF_ISTEREO) sum.n_ist += (uint32_t)n_gr;
        }
        prev[0] = f[0]; prev[1] = f[1]; prev[2] = f[2]; prev[3] = f[3];
        p += (uint32_t)fs;
        if (!last) cur = nxt;
    }
    sum.tail_pos = p;
    a.summaries[s] = sum;
}

/* stream of frame number t (frames of the taken streams counted through): last stream whose frame_base <= t */
__device__ __forceinline__ uint32_t stream_of_frame(const StreamBase *__restrict__ bases, uint32_t n_streams, uint32_t t)
{
    uint32_t lo = 0, hi = n_streams;
    while (hi - lo > 1)
    {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(&bases[mid].frame_base) <= t) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) k_stream_emit(EmitArgs a)
{
    const uint32_t t = blockIdx.x*blockDim.x + threadIdx.x;
    if (t >= a.n_frames) return;
    const uint32_t s = stream_of_frame(a.bases, a.n_streams, t);
    const ScanStream st = a.streams[s];
    const StreamBase sb = a.bases[s];
    const ScanFrame f = a.frames[st.slot_base + (t - sb.frame_base)];
    if (!(f.flags & SCANF_DECODABLE)) return;
    const uint8_t *src = a.raw + st.src_off + f.pos;
    const int per_frame = sb.n_gr*sb.nch;
    FrameRec r;
    r.md_bit = (sb.md_base + (uint64_t)f.md_before - (uint64_t)f.main_data_begin)*8u;
    r.gc_first = sb.gc_base + f.dec_index*(uint32_t)per_frame;
    for (int i = 0; i < 4; i++) r.hdr[i] = __ldg(src + i);
    r.nch = sb.nch;
    r.n_gr = sb.n_gr;
    for (int i = 0; i < 6; i++) r.pad[i] = 0;
    const int si_off = 4 + (hd_is_crc(r.hdr) ? 2 : 0);
    const int room = (int)f.frame_size - si_off;
    for (int i = 0; i < 32; i++) r.si[i] = i < room ? __ldg(src + si_off + i) : (uint8_t)0;
    for (int i = 0; i < 8; i++) r.pad2[i] = 0;
    a.frame_recs[sb.fr_base + f.dec_index] = r;
    for (int j = 0; j < per_frame; j++) a.order[sb.gc_base + f.dec_index*(uint32_t)per_frame + j - sb.gc_pad_before] = r.gc_first + (uint32_t)j;
    if (f.flags & SCANF_ISTEREO)
        for (int g = 0; g < sb.n_gr; g++) a.ist_list[sb.ist_base + f.ist_before + (uint32_t)g] = r.gc_first + (uint32_t)(g*sb.nch);
}

/* the payload of every frame -- decodable or not: the reservoir keeps undecodable frames' bytes too -- appended to
 * its stream's main-data run; one warp per frame, byte-wise (unaligned on both sides), coalesced across the warp */
__global__ void __launch_bounds__(256) k_payload_gather(EmitArgs a)
{
    const uint32_t t = (blockIdx.x*blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= a.n_frames) return;
    const uint32_t s = stream_of_frame(a.bases, a.n_streams, t);
    const ScanStream st = a.streams[s];
    const StreamBase sb = a.bases[s];
    const ScanFrame f = a.frames[st.slot_base + (t - sb.frame_base)];
    const uint8_t *src = a.raw + st.src_off + f.pos + f.payload_off;
    uint8_t *dst = a.md + sb.md_base + f.md_before;
    const int n = (int)f.frame_size - (int)f.payload_off;
    for (int i = lane; i < n; i += 32) dst[i] = __ldg(src + i);
}

/* tiles of a regular stream: granule g of the stream is descriptors gc_base + g nch .., every TG granules a tile,
 * halo = the two granules in front (none for the first tile: the decoder starts from zero state) */
__global__ void __launch_bounds__(256) k_stream_tiles(EmitArgs a)
{
    const uint32_t t = blockIdx.x*blockDim.x + threadIdx.x;
    if (t >= a.n_tiles) return;
    /* stream of tile t: last stream whose tile_base <= t */
    uint32_t lo = 0, hi = a.n_streams;
    while (hi - lo > 1)
    {
        const uint32_t mid = (lo + hi) >> 1;
        if (a.bases[mid].tile_base <= t) lo = mid; else hi = mid;
    }
    /* streams without tiles share their neighbour's tile_base: step to the one that owns tile t */
    while (lo + 1 < a.n_streams && a.bases[lo + 1].tile_base <= t) lo++;
    const StreamBase sb = a.bases[lo];
    const uint32_t tg = sb.nch == 1 ? MP3B_TILE_GRANULES_MONO : MP3B_TILE_GRANULES;
    const uint32_t local = t - sb.tile_base, n_granules = sb.n_dec*sb.n_gr;
    const uint32_t g0 = local*tg;
    TileDesc d;
    d.gc_first = sb.gc_base + g0*sb.nch;
    d.n_gr = (uint8_t)min(tg, n_granules - g0);
    d.nch = sb.nch;
    d.pad[0] = d.pad[1] = 0;
    d.pcm_off = sb.pcm_base + (uint64_t)g0*576u*sb.nch;
    for (int c = 0; c < 2; c++)
        for (int h = 0; h < 2; h++)
            d.halo[c][h] = (c < sb.nch && g0 > (uint32_t)h) ? (int32_t)(d.gc_first - (uint32_t)(h + 1)*sb.nch + (uint32_t)c) : -1;
    a.tiles[t] = d;
}

}  // namespace

cudaError_t launch_stream_scan(const ScanArgs &a, cudaStream_t s)
{
    if (!a.n_streams) return cudaSuccess;
    k_stream_scan<<<(a.n_streams + 63)/64, 64, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_stream_emit(const EmitArgs &a, cudaStream_t s)
{
    if (a.n_frames)
    {
        k_stream_emit<<<(a.n_frames + 255)/256, 256, 0, s>>>(a);
        k_payload_gather<<<(uint32_t)(((uint64_t)a.n_frames*32 + 255)/256), 256, 0, s>>>(a);
    }
    if (a.n_tiles) k_stream_tiles<<<(a.n_tiles + 255)/256, 256, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace mp3b
