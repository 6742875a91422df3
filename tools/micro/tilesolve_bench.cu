// Register-tiled Gauss-Jordan (hs_tilesolve.cuh) in the setting of the fused assembly kernels: resident CTAs of SPC
// systems x NW warps, the systems staged in shared memory; cycles per elimination step (clock64 around the solve) and
// ns per system with the SM full; a sample is checked against a host solve.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/tilesolve_bench tools/micro/tilesolve_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <complex>
#include <cuda_runtime.h>
#include "../../hydroscatt_b200/csrc/hs_staged.cuh"
using namespace hs;

template <int ROWS, int RT, int CT, int NW, int SPC, int MINB, int OLD>
__global__ void __launch_bounds__(32 * NW * SPC, MINB) k_bench(double2 *sys, int n, int ngroups, int *cursor, int *sing, unsigned long long *cyc)
{
    typedef TileSolve<ROWS, RT, CT, NW> TS;
    extern __shared__ __align__(16) double2 sm[];   // [SPC][n][2n]
    __shared__ typename TS::Buf bufs[SPC];
    __shared__ unsigned long long keyslot[SPC][2];
    __shared__ int s_next, s_sing;
    const int sg = threadIdx.x / (32 * NW), t = threadIdx.x % (32 * NW);
    const int per = n * 2 * n;
    for (;;) {
        if (threadIdx.x == 0) s_next = atomicAdd(cursor, 1);
        __syncthreads();
        const int grp = s_next;
        if (grp >= ngroups) return;
        double2 *g = sys + (size_t)grp * SPC * per;
        for (int i = threadIdx.x; i < SPC * per; i += blockDim.x) sm[i] = g[i];
        __syncthreads();
        SysDesc d;
        d.base = sm + sg * per; d.n = n; d.ld = 2 * n; d.off = 0; d.stride = 1; d.rhs = n;
        const long long t0 = clock64();
        if (OLD) solve_grp2s(d, t, 32 * NW, 1 + sg, keyslot[sg], &s_sing);
        else TS::solve(d, t, 1 + sg, &bufs[sg], &s_sing);
        const long long t1 = clock64();
        if (t == 0) atomicAdd(cyc, (unsigned long long)(t1 - t0));
        __syncthreads();
        for (int i = threadIdx.x; i < SPC * per; i += blockDim.x) g[i] = sm[i];
        __syncthreads();
    }
}

typedef std::complex<double> cd_t;
static void host_solve(const cd_t *in, int n, std::vector<cd_t> &X)
{
    std::vector<cd_t> a(in, in + (size_t)n * 2 * n);
    const int ld = 2 * n;
    for (int k = 0; k < n; k++) {
        int p = k; double best = -1;
        for (int r = k; r < n; r++) { double m = fabs(a[r * ld + k].real()) + fabs(a[r * ld + k].imag()); if (m > best) { best = m; p = r; } }
        for (int j = 0; j < ld; j++) std::swap(a[k * ld + j], a[p * ld + j]);
        const cd_t inv = 1.0 / a[k * ld + k];
        for (int j = 0; j < ld; j++) a[k * ld + j] *= inv;
        for (int r = 0; r < n; r++) if (r != k) { const cd_t f = a[r * ld + k]; for (int j = 0; j < ld; j++) a[r * ld + j] -= f * a[k * ld + j]; }
    }
    X.resize((size_t)n * n);
    for (int r = 0; r < n; r++) for (int j = 0; j < n; j++) X[r * n + j] = a[r * ld + n + j];
}

template <int ROWS, int RT, int CT, int NW, int SPC, int MINB, int OLD>
static void run(int n, int ngroups, int sms)
{
    const size_t per = (size_t)n * 2 * n;
    const int nsys = ngroups * SPC;
    std::vector<cd_t> h(per * nsys);
    srand(1234 + n);
    for (auto &x : h) x = cd_t(rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5);
    double2 *d; int *cur, *sing; unsigned long long *cyc;
    cudaMalloc(&d, per * nsys * 16); cudaMalloc(&cur, 4); cudaMalloc(&sing, 4); cudaMalloc(&cyc, 8);
    auto kern = k_bench<ROWS, RT, CT, NW, SPC, MINB, OLD>;
    const size_t smem = (size_t)SPC * per * 16;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * NW * SPC, smem);
    if (occ > MINB) occ = MINB;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f; unsigned long long hc = 0;
    for (int rep = 0; rep < 3; rep++) {
        cudaMemcpy(d, h.data(), per * nsys * 16, cudaMemcpyHostToDevice);
        cudaMemset(cur, 0, 4); cudaMemset(sing, 0, 4); cudaMemset(cyc, 0, 8);
        cudaEventRecord(e0);
        kern<<<sms * occ, 32 * NW * SPC, smem>>>(d, n, ngroups, cur, sing, cyc);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
        cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
    }
    cudaError_t e = cudaGetLastError();
    std::vector<cd_t> out(per * 8);
    cudaMemcpy(out.data(), d, per * 8 * 16, cudaMemcpyDeviceToHost);
    double err = 0;
    for (int s = 0; s < 8; s++) {
        std::vector<cd_t> X;
        host_solve(&h[s * per], n, X);
        double sc = 0;
        for (auto &x : X) sc = std::max(sc, std::abs(x));
        for (int r = 0; r < n; r++) for (int j = 0; j < n; j++) err = std::max(err, std::abs(out[s * per + (size_t)r * 2 * n + n + j] - X[r * n + j]) / sc);
    }
    printf("%s<%2d,%d,%d,%d> x%d  n %2d  regs %3d  %d CTAs/SM (%2d systems/SM)  %7.3f ms  %7.1f ns/system  %6.0f cycles/step in the CTA  %6.1f SM-cycles/(system step)  err %.1e %s\n",
           OLD ? "grp2s " : "tile  ", ROWS, RT, CT, NW, SPC, n, fa.numRegs, occ, occ * SPC, best, best * 1e6 / nsys, (double)hc / ((double)nsys * n),
           best * 1e-3 * 1.965e9 * sms / ((double)nsys * n), err, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d); cudaFree(cur); cudaFree(sing); cudaFree(cyc);
}

int main(int argc, char **argv)
{
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int sms = pr.multiProcessorCount;
    const int N = argc > 1 ? atoi(argv[1]) : 20000;  // groups of systems (small for compute-sanitizer --tool racecheck)
    // wide m >= 1: two systems of n <= 32 per 256-thread CTA, 2 CTAs per SM
    for (int n : {17, 20, 24, 32}) { run<32, 4, 4, 4, 2, 2, 0>(n, N, sms); run<32, 4, 4, 4, 2, 2, 1>(n, N, sms); }
    // narrow m >= 1: two systems of n <= 16 per 128-thread CTA, 4 CTAs per SM
    for (int n : {9, 12, 16}) { run<16, 2, 4, 2, 2, 4, 0>(n, 2 * N, sms); run<16, 2, 4, 2, 2, 4, 1>(n, 2 * N, sms); }
    // m = 0 half systems: wide (n <= 16, four per 256-thread CTA), narrow (n <= 8, four per 128-thread CTA)
    for (int n : {9, 11, 16}) { run<16, 2, 4, 2, 4, 2, 0>(n, N, sms); run<16, 2, 4, 2, 4, 2, 1>(n, N, sms); }
    for (int n : {5, 6, 8}) { run<8, 1, 4, 1, 4, 4, 0>(n, 2 * N, sms); run<8, 1, 4, 1, 4, 4, 1>(n, 2 * N, sms); }
    return 0;
}
