// Throughput of the register-resident Gauss-Jordan (hs_regsolve.cuh) with many systems in flight per SM: resident CTAs
// fetch systems [A | B] (n x 2n complex, L2-resident) from a queue, solve them in registers and write A^-1 B back.
// Reports ns per system and SM cycles per (system, elimination step); checks a sample against a host solve.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/regsolve_bench tools/micro/regsolve_bench.cu && /tmp/regsolve_bench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <complex>
#include <cuda_runtime.h>
#include "../../hydroscatt_b200/csrc/hs_staged.cuh"
using namespace hs;

template <int ROWS, int CPL, int NW, int SPC, int MINB>
__global__ void __launch_bounds__(32 * NW * SPC, MINB) k_bench(double2 *sys, int n, int nsys, int *cursor, int *sing)
{
    __shared__ double2 colbuf[SPC][2 * ROWS];
    __shared__ int s_next[SPC];
    const int sg = threadIdx.x / (32 * NW), t = threadIdx.x % (32 * NW), w = t >> 5, lane = t & 31;
    RegSolve<ROWS, CPL, NW> R;
    for (;;) {
        int s;
        if (NW > 1) {
            if (t == 0) s_next[sg] = atomicAdd(cursor, 1);
            asm volatile("bar.sync %0, %1;" ::"r"(1 + sg), "r"(32 * NW) : "memory");
            s = s_next[sg];
            asm volatile("bar.sync %0, %1;" ::"r"(1 + sg), "r"(32 * NW) : "memory");
        } else {
            s = 0;
            if (lane == 0) s = atomicAdd(cursor, 1);
            s = __shfl_sync(0xffffffffu, s, 0);
        }
        if (s >= nsys) return;
        SysDesc d;
        d.base = sys + (size_t)s * n * 2 * n; d.n = n; d.ld = 2 * n; d.off = 0; d.stride = 1; d.rhs = n;
        R.load(d, w, lane, [](const double2 *p) { return __ldg(p); });
        R.eliminate(n, w, lane, 1 + sg, colbuf[sg], sing);
        R.store(n, w, lane, [&](int row, int col, double2 x) { d.base[(size_t)row * d.ld + n + col] = x; });
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

template <int ROWS, int CPL, int NW, int SPC, int MINB>
static void run(int n, int nsys, int sms)
{
    const size_t per = (size_t)n * 2 * n;
    std::vector<cd_t> h(per * nsys);
    srand(1234 + n);
    for (auto &x : h) x = cd_t(rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5);
    double2 *d; int *cur, *sing;
    cudaMalloc(&d, per * nsys * 16); cudaMalloc(&cur, 4); cudaMalloc(&sing, 4);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_bench<ROWS, CPL, NW, SPC, MINB>, 32 * NW * SPC, 0);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_bench<ROWS, CPL, NW, SPC, MINB>);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 4; rep++) {
        cudaMemcpy(d, h.data(), per * nsys * 16, cudaMemcpyHostToDevice);
        cudaMemset(cur, 0, 4); cudaMemset(sing, 0, 4);
        cudaEventRecord(e0);
        k_bench<ROWS, CPL, NW, SPC, MINB><<<sms * occ, 32 * NW * SPC>>>(d, n, nsys, cur, sing);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
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
    const double cyc = best * 1e-3 * 1.965e9 * sms / ((double)nsys * n);
    printf("<%2d,%2d,%d> x%d  n %2d  regs %3d  %d CTAs/SM (%2d systems/SM)  %8.3f ms  %7.1f ns/system  %6.1f SM-cycles/(system step)  err %.1e  %s\n",
           ROWS, CPL, NW, SPC, n, fa.numRegs, occ, occ * SPC, best, best * 1e6 / nsys, cyc, err, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d); cudaFree(cur); cudaFree(sing);
}

int main()
{
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int sms = pr.multiProcessorCount;
    const int N = 60000;
    for (int n : {18, 20, 24, 28, 32}) run<32, 16, 4, 2, 2>(n, N, sms);
    for (int n : {20, 32}) run<32, 16, 4, 2, 3>(n, N, sms);
    for (int n : {20, 32}) run<32, 16, 4, 1, 6>(n, N, sms);
    for (int n : {20, 32}) run<32, 8, 8, 1, 4>(n, N, sms);
    for (int n : {9, 11, 12, 14, 16}) run<16, 16, 1, 4, 4>(n, 2 * N, sms);
    for (int n : {12, 16}) run<16, 16, 1, 4, 5>(n, 2 * N, sms);
    for (int n : {12, 16}) run<16, 8, 2, 2, 8>(n, 2 * N, sms);
    for (int n : {3, 4, 6, 8}) run<8, 4, 1, 8, 4>(n, 4 * N, sms);
    for (int n : {6, 8}) run<8, 4, 1, 8, 8>(n, 4 * N, sms);
    return 0;
}
