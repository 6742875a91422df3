// Microbenchmark: FP64 DFMA vs DMMA (mma.sync.m8n8k4.f64) peak on one GPU, alone and mixed.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, int iters)
{
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = i;
    double f[8];
#pragma unroll
    for (int i = 0; i < 8; i++) f[i] = i;
    for (int it = 0; it < iters; it++) {
        if (MODE & 1) {
#pragma unroll
            for (int i = 0; i < 8; i++) dmma(c[2 * i], c[2 * i + 1], a, b);
        }
        if (MODE & 2) {
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) f[i] = fma(f[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i];
#pragma unroll
    for (int i = 0; i < 8; i++) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
double run(double *d, int ctas, int iters)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<ctas, 256>>>(d, 10);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<ctas, 256>>>(d, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    // per iteration per warp: DMMA 8 x (8*8*4) FMA = 2048 FMA; DFMA 64 x 32 = 2048 FMA
    double fma_per_warp_iter = ((MODE & 1) ? 2048.0 : 0.0) + ((MODE & 2) ? 2048.0 : 0.0);
    double flops = 2.0 * fma_per_warp_iter * iters * (double)ctas * 8;
    return flops / (ms * 1e-3) / 1e12;
}
int main()
{
    double *d;
    cudaMalloc(&d, 148 * 8 * 256 * 8 * 8);
    for (int per = 1; per <= 8; per *= 2) {
        int ctas = 148 * per;
        printf("CTAs/SM %d: DMMA %.2f TF  DFMA %.2f TF  mixed %.2f TF\n", per, run<1>(d, ctas, 20000), run<2>(d, ctas, 20000), run<3>(d, ctas, 20000));
    }
    return 0;
}
