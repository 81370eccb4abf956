/* Microbenchmark: dependent-issue latency of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a, and how many
 * independent chains a warp needs to keep the FP32 pipe busy.  One warp per SM quadrant (4 warps per SM), CHAINS
 * independent accumulators each, clock64() around the loop.
 *   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/lat tools/mb_ffma_latency.cu && /tmp/lat
 * Why it matters: the window FIR of k_l3_transform is 8 dependent FFMA2 per output sample and chain. */
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int CHAINS>
__global__ void k(float *out, long long *cycles, int iters, float a)
{
    float x[CHAINS];
    unsigned long long p[CHAINS];
    for (int i = 0; i < CHAINS; i++)
    {
        x[i] = threadIdx.x*0.001f + i;
        p[i] = ((unsigned long long)__float_as_uint(x[i]) << 32) | __float_as_uint(x[i] + 0.5f);
    }
    const unsigned long long c2 = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(a);
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int r = 0; r < 16; r++)
#pragma unroll
            for (int i = 0; i < CHAINS; i++)
            {
                if (MODE == 0) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(x[i]) : "f"(x[i]), "f"(a), "f"(a));
                else asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p[i]) : "l"(p[i]), "l"(c2), "l"(c2));
            }
    }
    const long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < CHAINS; i++) s += x[i] + __uint_as_float((unsigned)p[i]);
    out[blockIdx.x*blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE, int CHAINS>
void run(float *d, long long *dc, int warps_per_sm)
{
    const int iters = 2000;
    long long c = 0;
    for (int rep = 0; rep < 2; rep++) k<MODE, CHAINS><<<148, warps_per_sm*32>>>(d, dc, iters, 0.999f);
    cudaMemcpy(&c, dc, sizeof(c), cudaMemcpyDeviceToHost);
    const double per_instr = (double)c/((double)iters*16*CHAINS);
    printf("%s  warps/SM %2d  chains %d: %.2f clk per instruction per warp  -> %s %.1f clk  (pipe share per scheduler %.0f%%)\n",
           MODE ? "FFMA2" : "FFMA ", warps_per_sm, CHAINS, per_instr, CHAINS == 1 ? "dependent latency" : "chain step every",
           per_instr*CHAINS, 100.0*(MODE ? 2.0 : 1.0)*(warps_per_sm/4.0)/per_instr);
}

int main()
{
    float *d;
    long long *dc;
    cudaMalloc(&d, 148*1024*4);
    cudaMalloc(&dc, 8);
    run<0, 1>(d, dc, 4); run<1, 1>(d, dc, 4);
    run<0, 2>(d, dc, 4); run<1, 2>(d, dc, 4);
    run<0, 4>(d, dc, 4); run<1, 4>(d, dc, 4);
    run<0, 8>(d, dc, 4); run<1, 8>(d, dc, 4);
    run<1, 2>(d, dc, 8); run<1, 4>(d, dc, 8); run<1, 2>(d, dc, 24); run<1, 4>(d, dc, 24);
    return 0;
}
