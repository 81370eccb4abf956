/*
 * tools/mb_dct_tcgen05.cu -- the north star's conditional, measured: the 32-point DCT-II of the synthesis filterbank
 * (k_l3_transform phase 4) as a tcgen05 3xTF32 GEMM against the FFMA2 Lee recursion the kernel uses.
 *
 *   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o /tmp/dct tools/mb_dct_tcgen05.cu && /tmp/dct
 *
 * Both variants transform tiles of 128 rows x 32 subbands that stay on chip: every thread owns one row in registers,
 * transforms it REPS times (output fed back, scaled to stay finite) and stores it once, so the timing is the
 * transform itself, not HBM.
 *   lee     Lee recursion in registers, half-size transforms paired in packed FP32 (the product's dct2_32_paired)
 *   tcgen05 D[128 x 32] = A[128 x 32] . C^T, C[n][k] = cos(pi (2k+1) n / 64), as three TF32 products
 *           (A_hi C_hi + A_lo C_hi + A_hi C_lo, accumulated in TMEM): per round every thread splits its row into
 *           tf32 hi / lo parts (cvt.rna.tf32 + one subtraction), writes both to shared memory in the no-swizzle K-major
 *           core-matrix layout (8 rows x 16 bytes), one elected thread issues 12 tcgen05.mma (4 K-steps x 3 products,
 *           M 128, N 32, K 8), completion through tcgen05.commit -> mbarrier, result back through tcgen05.ld (32x32b.x32).
 * Prints time per row for both and the largest deviation of each from a double-precision DCT.
 */
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__constant__ float c_sec[31];
__constant__ unsigned long long c_sec2[16];

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) { unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) { unsigned long long r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) { unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

template <int N> struct Lee2
{
    static __device__ __forceinline__ void run(unsigned long long *x, const unsigned long long *sec)
    {
        unsigned long long a[N/2], b[N/2];
#pragma unroll
        for (int i = 0; i < N/2; i++) { a[i] = add2(x[i], x[N - 1 - i]); b[i] = mul2(sub2(x[i], x[N - 1 - i]), sec[i]); }
        Lee2<N/2>::run(a, sec + N/2);
        Lee2<N/2>::run(b, sec + N/2);
#pragma unroll
        for (int k = 0; k < N/2; k++) { x[2*k] = a[k]; x[2*k + 1] = (k + 1 < N/2) ? add2(b[k], b[k + 1]) : b[k]; }
    }
};
template <> struct Lee2<1> { static __device__ __forceinline__ void run(unsigned long long *, const unsigned long long *) {} };

__device__ __forceinline__ void dct32_lee(float *x)
{
    unsigned long long p[16];
#pragma unroll
    for (int i = 0; i < 16; i++) p[i] = pack2(x[i] + x[31 - i], (x[i] - x[31 - i])*c_sec[i]);
    Lee2<16>::run(p, c_sec2);
    float a[16], b[16];
#pragma unroll
    for (int k = 0; k < 16; k++) unpack2(p[k], a[k], b[k]);
#pragma unroll
    for (int k = 0; k < 16; k++) { x[2*k] = a[k]; x[2*k + 1] = k + 1 < 16 ? b[k] + b[k + 1] : b[k]; }
}

constexpr float kFeedback = 0.125f;       /* keeps the fed-back rows bounded (the DCT gains up to 32) */

__global__ void __launch_bounds__(128) k_lee(const float *__restrict__ in, float *__restrict__ out, int n_tiles, int reps)
{
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
    {
        const size_t row = (size_t)tile*128 + threadIdx.x;
        float x[32];
#pragma unroll
        for (int k = 0; k < 8; k++) { const float4 q = reinterpret_cast<const float4 *>(in + row*32)[k]; x[4*k] = q.x; x[4*k + 1] = q.y; x[4*k + 2] = q.z; x[4*k + 3] = q.w; }
        for (int r = 0; r < reps; r++)
        {
            dct32_lee(x);
            if (r + 1 < reps)
            {
#pragma unroll
                for (int k = 0; k < 32; k++) x[k] *= kFeedback;
            }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) reinterpret_cast<float4 *>(out + row*32)[k] = make_float4(x[4*k], x[4*k + 1], x[4*k + 2], x[4*k + 3]);
    }
}

/* ---- tcgen05 path ---- */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float to_tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return __uint_as_float(r); }

/* K-major, no swizzle: core matrix = 8 rows x 16 bytes, contiguous; next 16 bytes of K: +LBO; next 8 rows: +SBO */
constexpr uint32_t kLBO = 128, kSBO = 1024;
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(kLBO >> 4) << 16;
    d |= (uint64_t)(kSBO >> 4) << 32;
    d |= (uint64_t)1 << 46;                 /* descriptor version of sm_100 */
    return d;                               /* base offset 0, layout type 0 = no swizzle */
}
/* instruction descriptor, kind::tf32: D f32, A / B tf32, both K-major, N = 32, M = 128 */
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);

struct __align__(1024) SmemTC
{
    float a_hi[128*32], a_lo[128*32];       /* 16 KB each, core-matrix layout */
    float c_hi[32*32], c_lo[32*32];         /* 4 KB each */
    unsigned long long bar;
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(128) k_tc(const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ cmat_hi,
                                            const float *__restrict__ cmat_lo, int n_tiles, int reps)
{
    extern __shared__ __align__(1024) unsigned char raw[];
    SmemTC &S = *reinterpret_cast<SmemTC *>(raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    /* C[n][k] in the same layout (n = row of the B operand) */
    for (int e = tid; e < 32*32; e += 128)
    {
        const int n = e >> 5, k = e & 31;
        const int off = (n >> 3)*256 + (k >> 2)*32 + (n & 7)*4 + (k & 3);       /* in floats: SBO 1024 B, LBO 128 B, row 16 B */
        S.c_hi[off] = cmat_hi[e];
        S.c_lo[off] = cmat_lo[e];
    }
    if (tid == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&S.bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" :: "r"(smem_u32(&S.tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = S.tmem_base;
    uint32_t parity = 0;
    const uint32_t a_hi = smem_u32(S.a_hi), a_lo = smem_u32(S.a_lo), c_hi = smem_u32(S.c_hi), c_lo = smem_u32(S.c_lo);
    float *my_hi = S.a_hi + (tid >> 3)*256 + (tid & 7)*4, *my_lo = S.a_lo + (tid >> 3)*256 + (tid & 7)*4;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
    {
        const size_t row = (size_t)tile*128 + tid;
        float x[32];
#pragma unroll
        for (int k = 0; k < 8; k++) { const float4 q = reinterpret_cast<const float4 *>(in + row*32)[k]; x[4*k] = q.x; x[4*k + 1] = q.y; x[4*k + 2] = q.z; x[4*k + 3] = q.w; }
        for (int r = 0; r < reps; r++)
        {
            /* split the row, 16 bytes of K at a time, into the two operand tiles */
#pragma unroll
            for (int c = 0; c < 8; c++)
            {
                float4 h, l;
                h.x = to_tf32(x[4*c]);     l.x = to_tf32(x[4*c] - h.x);
                h.y = to_tf32(x[4*c + 1]); l.y = to_tf32(x[4*c + 1] - h.y);
                h.z = to_tf32(x[4*c + 2]); l.z = to_tf32(x[4*c + 2] - h.z);
                h.w = to_tf32(x[4*c + 3]); l.w = to_tf32(x[4*c + 3] - h.w);
                *reinterpret_cast<float4 *>(my_hi + c*32) = h;
                *reinterpret_cast<float4 *>(my_lo + c*32) = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       /* generic writes -> tensor-core reads */
            __syncthreads();
            if (tid == 0)
            {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                int first = 1;
#pragma unroll
                for (int prod = 0; prod < 3; prod++)
                {
                    const uint32_t a_base = prod == 1 ? a_lo : a_hi, b_base = prod == 2 ? c_lo : c_hi;
#pragma unroll
                    for (int kk = 0; kk < 4; kk++)
                    {
                        const uint64_t da = make_desc(a_base + kk*2*kLBO), db = make_desc(b_base + kk*2*kLBO);
                        const uint32_t acc = first ? 0u : 1u;
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                                     :: "r"(tmem), "l"(da), "l"(db), "r"(kIdesc), "r"(acc) : "memory");
                        first = 0;
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&S.bar)) : "memory");
            }
            {
                uint32_t done;
                do
                {
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(done) : "r"(smem_u32(&S.bar)), "r"(parity) : "memory");
                } while (!done);
                parity ^= 1;
            }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t v[32];
            const uint32_t taddr = tmem + ((uint32_t)(warp*32) << 16);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                         "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                         "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                           "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                           "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const float scale = r + 1 < reps ? kFeedback : 1.f;
#pragma unroll
            for (int k = 0; k < 32; k++) x[k] = __uint_as_float(v[k])*scale;
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();          /* everybody has read D before the next round's first MMA overwrites it */
        }
#pragma unroll
        for (int k = 0; k < 8; k++) reinterpret_cast<float4 *>(out + row*32)[k] = make_float4(x[4*k], x[4*k + 1], x[4*k + 2], x[4*k + 3]);
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" :: "r"(tmem) : "memory");
}

static float tf32_round_host(float x)
{
    uint32_t u;
    memcpy(&u, &x, 4);
    u += 0x1000u;             /* round to nearest, ties away: what cvt.rna does on the magnitude */
    u &= 0xFFFFE000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}

int main(int argc, char **argv)
{
    const int n_tiles = argc > 1 ? atoi(argv[1]) : 148*64, reps = argc > 2 ? atoi(argv[2]) : 64;
    const size_t n_rows = (size_t)n_tiles*128;
    std::vector<float> h_in(n_rows*32), h_a(n_rows*32), h_b(n_rows*32);
    uint32_t lcg = 7;
    for (auto &v : h_in) { lcg = lcg*1664525u + 1013904223u; v = ((lcg >> 8)/16777216.0f*2.f - 1.f)*1.0e4f; }
    float sec[31], *p = sec;
    for (int n = 32; n >= 2; n >>= 1)
        for (int i = 0; i < n/2; i++) *p++ = (float)(0.5/cos(M_PI*(2*i + 1)/(2.0*n)));
    unsigned long long sec2[16] = { 0 };
    for (int i = 0; i < 15; i++) { uint32_t u; memcpy(&u, &sec[16 + i], 4); sec2[i] = ((unsigned long long)u << 32) | u; }
    CK(cudaMemcpyToSymbol(c_sec, sec, sizeof(sec)));
    CK(cudaMemcpyToSymbol(c_sec2, sec2, sizeof(sec2)));
    std::vector<float> chi(1024), clo(1024);
    for (int n = 0; n < 32; n++)
        for (int k = 0; k < 32; k++)
        {
            const float c = (float)cos(M_PI*(2*k + 1)*n/64.0);
            chi[n*32 + k] = tf32_round_host(c);
            clo[n*32 + k] = tf32_round_host(c - chi[n*32 + k]);
        }
    float *d_in, *d_out, *d_chi, *d_clo;
    CK(cudaMalloc(&d_in, n_rows*32*4)); CK(cudaMalloc(&d_out, n_rows*32*4)); CK(cudaMalloc(&d_chi, 4096)); CK(cudaMalloc(&d_clo, 4096));
    CK(cudaMemcpy(d_in, h_in.data(), n_rows*32*4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_chi, chi.data(), 4096, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_clo, clo.data(), 4096, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(k_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemTC) + 1024));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int grid = 148*4;
    auto time_ms = [&](int which, int r) {
        float best = 1e30f;
        for (int it = 0; it < 4; it++)
        {
            CK(cudaEventRecord(e0));
            if (which == 0) k_lee<<<grid, 128>>>(d_in, d_out, n_tiles, r);
            else k_tc<<<grid, 128, sizeof(SmemTC) + 1024>>>(d_in, d_out, d_chi, d_clo, n_tiles, r);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (it) best = ms < best ? ms : best;
        }
        return best;
    };
    /* accuracy: one transform */
    k_lee<<<grid, 128>>>(d_in, d_out, n_tiles, 1);
    CK(cudaMemcpy(h_a.data(), d_out, n_rows*32*4, cudaMemcpyDeviceToHost));
    k_tc<<<grid, 128, sizeof(SmemTC) + 1024>>>(d_in, d_out, d_chi, d_clo, n_tiles, 1);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h_b.data(), d_out, n_rows*32*4, cudaMemcpyDeviceToHost));
    double ea = 0, eb = 0, mx = 0;
    for (size_t r = 0; r < 2048 && r < n_rows; r++)
        for (int n = 0; n < 32; n++)
        {
            double s = 0;
            for (int k = 0; k < 32; k++) s += (double)h_in[r*32 + k]*cos(M_PI*(2*k + 1)*n/64.0);
            ea = fmax(ea, fabs(h_a[r*32 + n] - s)); eb = fmax(eb, fabs(h_b[r*32 + n] - s)); mx = fmax(mx, fabs(s));
        }
    printf("accuracy vs double (|y| up to %.3g): lee max err %.3g (%.2e relative), tcgen05 3xTF32 max err %.3g (%.2e relative)\n",
           mx, ea, ea/mx, eb, eb/mx);
    for (int which = 0; which < 2; which++)
    {
        const float t1 = time_ms(which, 1), tr = time_ms(which, reps);
        const double per_row_ns = (double)(tr - t1)*1e6/((double)(reps - 1)*(double)n_rows);
        printf("%-8s %d tiles x %d rounds: %.3f ms (1 round %.3f ms) -> %.4f ns per row on the whole GPU = %.1f G rows/s = %.0f Gsamples/s of synthesis input\n",
               which ? "tcgen05" : "lee", n_tiles, reps, tr, t1, per_row_ns, 1.0/per_row_ns, 32.0/per_row_ns);
    }
    return 0;
}
