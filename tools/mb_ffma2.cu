/* Microbenchmark behind DESIGN.md section 3: issue rate of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
 *   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ffma tools/mb_ffma2.cu && /tmp/ffma
 * Measured on B200: FFMA 3.9 warp-instructions/clk/SM, FFMA2 2.0 -- the same 124-127 FMA lanes/clk/SM, i.e. FFMA2
 * does not add FLOPs, it halves the issue slots the same FLOPs cost. */
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float *out, int iters, float a)
{
    float x[8]; unsigned long long p[8];
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x*0.001f + i; p[i] = ((unsigned long long)__float_as_uint(x[i]) << 32) | __float_as_uint(x[i] + 0.5f); }
    unsigned long long c2 = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(a);
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < 8; i++)
            {
                if (MODE == 0) x[i] = fmaf(x[i], a, 0.25f);
                else asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p[i]) : "l"(p[i]), "l"(c2), "l"(c2));
            }
    }
    float s = 0; for (int i = 0; i < 8; i++) s += x[i] + __uint_as_float((unsigned)p[i]);
    out[blockIdx.x*blockDim.x + threadIdx.x] = s;
}
int main()
{
    float *d; cudaMalloc(&d, 148*8*256*4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps = 4; warps <= 32; warps *= 2)
      for (int mode = 0; mode < 2; mode++)
      {
        int threads = 256, blocks = 148*warps*32/threads, iters = 4000;
        for (int rep = 0; rep < 2; rep++)
        {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<blocks, threads>>>(d, iters, 0.999f); else k<1><<<blocks, threads>>>(d, iters, 0.999f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double inst = (double)blocks*threads/32*iters*64;   // warp-instructions
        double cyc = ms*1e-3*1.965e9;
        printf("warps/SM %2d mode %s: %.3f ms, warp-instr/clk/SM %.2f, FMA lanes/clk/SM %.1f\n", warps, mode ? "FFMA2" : "FFMA ", ms, inst/cyc/148, inst/cyc/148*32*(mode ? 2 : 1));
      }
    return 0;
}
