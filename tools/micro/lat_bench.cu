// Latencies on this GPU (cycles): dependent DFMA, double division, LDS chain, named barrier, shuffle, sqrt, one smem row update.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, double x0)
{
    __shared__ double2 sm[1024];
    __shared__ int idx[64];
    const int tid = threadIdx.x;
    for (int t = tid; t < 1024; t += blockDim.x) sm[t] = make_double2(1.0 + 1e-3 * t, 0.5);
    if (tid < 64) idx[tid] = (tid * 7 + 1) & 63;
    __syncthreads();
    const int N = 256;
    double x = x0 + tid * 1e-9, y = 1.000001;
    long long t0, t1;
    t0 = clock64();
    for (int i = 0; i < N; i++) x = fma(x, y, 1e-9);
    t1 = clock64(); if (tid == 0) cyc[0] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; i++) x = 1.0000001 / x + 0.5;
    t1 = clock64(); if (tid == 0) cyc[1] = (t1 - t0) / N;
    int p = tid & 63;
    t0 = clock64();
    for (int i = 0; i < N; i++) p = idx[p];
    t1 = clock64(); if (tid == 0) cyc[2] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; i++) asm volatile("bar.sync %0, %1;" ::"r"(1 + (tid >> 7)), "r"(128) : "memory");
    t1 = clock64(); if (tid == 0) cyc[3] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; i++) __syncthreads();
    t1 = clock64(); if (tid == 0) cyc[4] = (t1 - t0) / N;
    double z = x;
    t0 = clock64();
    for (int i = 0; i < N; i++) z = __shfl_sync(0xffffffffu, z, (tid + 1) & 31) + 1.0;
    t1 = clock64(); if (tid == 0) cyc[5] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; i++) x = sqrt(x) + 1.5;
    t1 = clock64(); if (tid == 0) cyc[6] = (t1 - t0) / N;
    double2 v = sm[tid];
    t0 = clock64();
    for (int i = 0; i < N; i++) { double2 f = sm[(tid + i) & 1023]; v.x -= f.x * y; v.y -= f.y * y; sm[(tid + i * 3) & 1023] = v; }
    t1 = clock64(); if (tid == 0) cyc[7] = (t1 - t0) / N;
    out[blockIdx.x * blockDim.x + tid] = x + p + z + v.x;
}
int main()
{
    double *out; long long *cyc, h[8];
    cudaMalloc(&out, 148 * 256 * 8); cudaMalloc(&cyc, 64);
    k<<<1, 256>>>(out, cyc, 1.3);
    cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    printf("1 CTA: dep DFMA %lld, dep DDIV(+add) %lld, LDS chase %lld, bar.sync(128) %lld, __syncthreads(256) %lld, shfl+add %lld, dsqrt+add %lld, LDS/FMA/STS row %lld  (%s)\n",
           h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], cudaGetErrorString(cudaGetLastError()));
    return 0;
}
