/*
 * k_order.cu -- lane order of the entropy kernel across the streams of a batch.
 *
 * The entropy kernel decodes 32 granule-channels per warp in lock step (one lane each), so a warp takes as long as
 * its longest lane.  The host walker orders every stream's granule-channels by size on its own (build_order in
 * host_parse.cpp).  Here the streams of a window of 32 consecutive streams are merged: inside a window the
 * granule-channels are sorted by (big_values / 2, part_23_length / 16) -- the two side-info fields L3_huffman's loops
 * run on (minimp3.h:742-877; read by L3_read_side_info, minimp3.h:484-607) -- and, among equals, by 256-slot stretch
 * of the stream timeline, stream after stream.  Streams that are alike -- same material,
 * same encoder settings; at best copies of each other, as in BASELINE.json's batch -- then fill whole warps with
 * lanes of equal length (lane efficiency 0.81 -> 0.99 on that batch).  A window of unrelated streams would lose a
 * few per cent against the per-stream order (the two fields predict the line count poorly across sample rates and
 * block types, and one stream per warp is the friendlier memory pattern), so a window is only merged when its
 * streams are alike: k_order_alike counts, per window, the slots whose two fields equal those of the same slot of
 * the window's first stream; below half, the window keeps the walker's order (stream, part_23_length / 16, slot).
 * The window is small because the lanes of a warp should read their bitstreams from neighbouring memory: with 32
 * random streams of a 1024-stream batch per warp the kernel takes twice as long (one bitstream load then touches
 * 32 pages instead of three).
 *
 * Two small kernels build a 64-bit key per granule-channel, a stable radix sort (CUB, part of the CUDA toolkit; a
 * preparation step like the H2D copies, once per plan) orders the list by it.
 */
#include <cuda_runtime.h>
#include <stdlib.h>

#include <cub/device/device_radix_sort.cuh>

#include "kernels.h"

namespace mp3b {

namespace {

constexpr int kWindowShift = 5;                   /* 32 streams per window */
constexpr int kSlotBlockShift = 8;                /* 256 slots per block of the slot field */

struct KeyLayout
{
    int pos_bits;       /* slot index inside the stream ... */
    int pos_shift;      /* ... without its lowest kSlotBlockShift bits */
    int total_bits;
};

/* stream of a granule-channel: last s with gc_base[s] <= gc */
__device__ __forceinline__ uint32_t stream_of(const uint32_t *__restrict__ gc_base, uint32_t n_streams, uint32_t gc)
{
    uint32_t lo = 0, hi = n_streams;
    while (hi - lo > 1)
    {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(gc_base + mid) <= gc) lo = mid; else hi = mid;
    }
    return lo;
}

/* part_23_length | big_values << 16 of a descriptor */
__device__ __forceinline__ uint32_t size_word(const GcDesc *__restrict__ gcs, uint32_t gc)
{
    return __ldg(reinterpret_cast<const uint32_t *>(gcs + gc) + 2);
}

/* stat[2 w] += slots of window w (outside its first stream) whose size fields equal the first stream's at the same
 * slot, stat[2 w + 1] += slots compared */
__global__ void __launch_bounds__(256) k_order_alike(const GcDesc *__restrict__ gcs, const uint32_t *__restrict__ order_in, uint32_t n,
                                                     const uint32_t *__restrict__ gc_base, uint32_t n_streams, uint32_t *__restrict__ stat)
{
    const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    uint32_t win = 0xFFFFFFFFu;
    bool counted = false, same = false;
    if (i < n)
    {
        const uint32_t gc = __ldg(order_in + i);
        const uint32_t s = stream_of(gc_base, n_streams, gc), s0 = s & ~((1u << kWindowShift) - 1);
        win = s >> kWindowShift;
        if (s != s0)
        {
            counted = true;
            const uint32_t slot = gc - __ldg(gc_base + s), first = __ldg(gc_base + s0);
            if (slot < __ldg(gc_base + s0 + 1) - first)
            {
                const uint32_t a = size_word(gcs, gc), b = size_word(gcs, first + slot);
                same = (a >> 16) == (b >> 16) && ((a & 0xffffu) >> 3) == ((b & 0xffffu) >> 3);
            }
        }
    }
    /* the list arrives stream by stream: a block almost always sits inside one window.  Warps of the block's first
     * window add up in shared memory and the block adds once (one atomic pair per warp on the same 64 words of HBM was
     * most of this kernel's time); anything else goes to the counters directly */
    __shared__ uint32_t s_cnt[2], s_win;
    if (threadIdx.x == 0) { s_cnt[0] = s_cnt[1] = 0; s_win = win; }
    __syncthreads();
    const uint32_t win0 = __shfl_sync(0xffffffffu, win, 0);
    if (__all_sync(0xffffffffu, win == win0))
    {
        const uint32_t n_same = __popc(__ballot_sync(0xffffffffu, same)), n_counted = __popc(__ballot_sync(0xffffffffu, counted));
        if ((threadIdx.x & 31) == 0 && win0 != 0xFFFFFFFFu && n_counted)
        {
            if (win0 == s_win)
            {
                atomicAdd(&s_cnt[0], n_same);
                atomicAdd(&s_cnt[1], n_counted);
            } else
            {
                atomicAdd(stat + 2*win0, n_same);
                atomicAdd(stat + 2*win0 + 1, n_counted);
            }
        }
    } else if (counted)
    {
        if (same) atomicAdd(stat + 2*win, 1u);
        atomicAdd(stat + 2*win + 1, 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt[1])
    {
        atomicAdd(stat + 2*s_win, s_cnt[0]);
        atomicAdd(stat + 2*s_win + 1, s_cnt[1]);
    }
}

/* key, most significant first, of a window whose streams are alike:
 *     window | big_values / 2 (8, descending) | part_23_length / 16 (8, descending) | slot / 256
 * and of any other window (the walker's per-stream order):
 *     window | stream inside the window (5) part_23_length / 16 (8) 0 (3) | slot / 256
 * The sort is stable and the list arrives in (stream, slot) order, so equal keys stay stream by stream, slot by slot:
 * the lanes of a warp are then granules of equal size from one stretch (256 slots, a second or two of audio, some
 * tens of KB of bitstream) of one stream, or of neighbouring streams where a stream runs out of them.  With the full
 * slot and the stream in the key (slot k of all 32 streams side by side) the same batches decoded 5-9 % slower and the
 * sort needed a fifth pass. */
__global__ void __launch_bounds__(256) k_order_keys(const GcDesc *__restrict__ gcs, const uint32_t *__restrict__ order_in, uint32_t n,
                                                    const uint32_t *__restrict__ gc_base, uint32_t n_streams, int pos_bits, int pos_shift,
                                                    const uint32_t *__restrict__ stat, int force, unsigned long long *__restrict__ keys)
{
    const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t gc = __ldg(order_in + i);
    const uint32_t lo = stream_of(gc_base, n_streams, gc);
    const uint32_t slot = gc - __ldg(gc_base + lo);
    const uint32_t w = size_word(gcs, gc);
    const uint32_t win = lo >> kWindowShift, in_win = lo & ((1u << kWindowShift) - 1);
    const bool alike = force >= 0 ? force != 0 : 2u*__ldg(stat + 2*win) >= __ldg(stat + 2*win + 1) && __ldg(stat + 2*win + 1) != 0;
    unsigned long long k = win;
    if (alike)
    {
        /* longest first: the entropy kernel's warps claim their groups in list order, so the groups claimed last --
         * the tail that decides when the kernel ends -- are the short ones */
        /* eight bits each (pairs of big values, 16 bits of main data): with the five window bits of a 1024-stream
         * batch and three for the slot stretch the key is 24 bits, three passes of the radix sort */
        const uint32_t len16 = min((w & 0xffffu) >> 4, 255u), big2 = min(w >> 17, 255u);
        k = k << 8 | (255u - big2);
        k = k << 8 | (255u - len16);
        k = k << pos_bits | (slot >> pos_shift);
    } else
    {
        k = k << 16 | (in_win << 11 | min((w & 0xffffu) >> 4, 255u) << 3);
        k = k << pos_bits | (slot >> pos_shift);
    }
    keys[i] = k;
}

int bits_for(uint64_t max_value)
{
    int b = 1;
    while (b < 63 && (max_value >> b)) b++;
    return b;
}

KeyLayout key_layout(uint32_t n_streams, uint32_t max_slots)
{
    KeyLayout L;
    const int slot_bits = bits_for(max_slots ? max_slots - 1 : 0);
    L.pos_shift = slot_bits < kSlotBlockShift ? slot_bits : kSlotBlockShift;
    L.pos_bits = slot_bits - L.pos_shift;
    const int fixed_bits = bits_for((n_streams - 1) >> kWindowShift) + 8 + 8;
    L.total_bits = fixed_bits + L.pos_bits;      /* <= 27 + 16 + 24 */
    return L;
}

}  // namespace

size_t order_sort_scratch_bytes(uint32_t n, uint32_t n_streams)
{
    size_t temp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n);
    /* two key arrays and the per-window counters in front of CUB's own scratch */
    return 2*((size_t)n*sizeof(unsigned long long) + 256) + ((((size_t)n_streams >> kWindowShift) + 1)*8 + 256) + temp + 256;
}

cudaError_t launch_order_sort(const GcDesc *gcs, const uint32_t *order_in, uint32_t n, uint32_t *order_out,
                              const uint32_t *gc_base, uint32_t n_streams, uint32_t max_slots,
                              void *scratch, size_t scratch_bytes, cudaStream_t s)
{
    if (!n || !n_streams) return cudaSuccess;
    const KeyLayout L = key_layout(n_streams, max_slots);
    const size_t key_bytes = ((size_t)n*sizeof(unsigned long long) + 255) & ~(size_t)255;
    const size_t stat_bytes = ((((size_t)n_streams >> kWindowShift) + 1)*8 + 255) & ~(size_t)255;
    unsigned long long *keys_in = (unsigned long long *)scratch;
    unsigned long long *keys_out = (unsigned long long *)((char *)scratch + key_bytes);
    uint32_t *stat = (uint32_t *)((char *)scratch + 2*key_bytes);
    void *temp = (char *)scratch + 2*key_bytes + stat_bytes;
    if (scratch_bytes < 2*key_bytes + stat_bytes) return cudaErrorInvalidValue;
    size_t temp_bytes = scratch_bytes - 2*key_bytes - stat_bytes;
    /* MP3B_DEBUG_ORDER_ALIKE=0 / 1: treat every window as unrelated / alike (experiments) */
    static const int force = [] { const char *p = getenv("MP3B_DEBUG_ORDER_ALIKE"); return p ? atoi(p) : -1; }();
    cudaError_t e = cudaMemsetAsync(stat, 0, stat_bytes, s);
    if (e != cudaSuccess) return e;
    if (force < 0) k_order_alike<<<(n + 255)/256, 256, 0, s>>>(gcs, order_in, n, gc_base, n_streams, stat);
    k_order_keys<<<(n + 255)/256, 256, 0, s>>>(gcs, order_in, n, gc_base, n_streams, L.pos_bits, L.pos_shift, stat, force, keys_in);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, order_in, order_out, (int)n, 0, L.total_bits, s);
}

}  // namespace mp3b
