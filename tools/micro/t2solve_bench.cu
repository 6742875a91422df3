// Cycles per elimination step of the two-warp register Gauss-Jordan of the two-layer kernel (t2_solve_rows), alone on an SM
// and with FP64-busy neighbour warps, with a per-section breakdown.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o t2solve_bench t2solve_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../hydroscatt_b200/csrc/hs_tm2l.cuh"
using namespace hs;

// copy of t2_solve_rows with clock reads between the sections (lane 0 of the L warp accumulates)
__device__ __forceinline__ void solve_rows_timed(const double2 *L, int ldl, const double2 *R, int ldr, double2 *out, int ldo, int NMq,
                                                 int part, int lane, int barid, double2 *colbuf, double2 *rowbuf, int *sing, long long *sec)
{
    const unsigned FULL = 0xffffffffu;
    double2 v[T2_NRANK];
    const bool realr = lane < NMq;
    const double2 *src = part ? R + lane * ldr : L + lane * ldl;
#pragma unroll
    for (int j = 0; j < T2_NRANK; j++) {
        double2 x = make_double2(0.0, 0.0);
        if (realr && j < NMq) x = src[j];
        else if (part == 0 && j == lane) x.x = 1.0;
        v[j] = x;
    }
    int vp = lane;
    double2 f = v[0];
    long long t0 = clock64();
    for (int k = 0; k < NMq; k++) {
        double2 *cb = colbuf + (k & 1) * T2_NRANK;
        if (part == 0 && lane < T2_NRANK) cb[lane] = f;
        asm volatile("bar.sync %0, %1;" ::"r"(barid), "r"(64) : "memory");
        double2 c = make_double2(0.0, 0.0);
        if (lane < T2_NRANK) c = cb[lane];
        { long long t = clock64(); sec[0] += t - t0; t0 = t; }
        const bool cand = lane < T2_NRANK && vp >= k;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(fabs(c.x) + fabs(c.y));
        const unsigned hi = (unsigned)(bits >> 32) + 1u, lo = ((unsigned)bits & 0xFFFFFF00u) | (unsigned)(255 - vp);
        const unsigned m1 = __reduce_max_sync(FULL, cand ? hi : 0u);
        const bool top = cand && hi == m1;
        const unsigned m2 = __reduce_max_sync(FULL, top ? lo : 0u);
        const int P = __ffs(__ballot_sync(FULL, top && lo == m2)) - 1;
        if (m1 == 1u && (m2 >> 8) == 0u && part == 0 && lane == 0) *sing = 1;
        { long long t = clock64(); sec[1] += t - t0; t0 = t; }
        if (lane == P) {
#pragma unroll
            for (int j = 0; j < T2_NRANK; j++) rowbuf[j] = v[j];
        }
        const double pvx = __shfl_sync(FULL, c.x, P), pvy = __shfl_sync(FULL, c.y, P);
        const int vpP = __shfl_sync(FULL, vp, P);
        const double rden = hs_rcp(pvx * pvx + pvy * pvy);
        const double2 inv = make_double2(pvx * rden, -pvy * rden);
        __syncwarp();
        if (lane < T2_NRANK) {
            const double2 raw = rowbuf[lane];
            rowbuf[T2_NRANK + lane] = make_double2(raw.x * inv.x - raw.y * inv.y, raw.x * inv.y + raw.y * inv.x);
        }
        __syncwarp();
        { long long t = clock64(); sec[2] += t - t0; t0 = t; }
        double2 fn = f;
        const double ncx = -c.x, ncy = -c.y;
#pragma unroll
        for (int j = 0; j < T2_NRANK; j++) {
            const double2 nk = rowbuf[T2_NRANK + j];
            double2 u = v[j];
            u.x = fma(c.y, nk.y, fma(ncx, nk.x, u.x));
            u.y = fma(ncy, nk.x, fma(ncx, nk.y, u.y));
            if (lane == P) u = nk;
            v[j] = u;
            if (j == k + 1) fn = u;
        }
        f = fn;
        if (vp == k) vp = vpP;
        if (lane == P) vp = k;
        { long long t = clock64(); sec[3] += t - t0; t0 = t; }
    }
    if (part == 1 && realr) {
#pragma unroll
        for (int j = 0; j < T2_NRANK; j++)
            if (j < NMq) out[vp * ldo + j] = v[j];
    }
}

// block: warps 0..3 = the solver group (2 classes x 2 parts); warps 4.. = FP64-busy neighbours (nbusy warps)
__global__ void k_bench(const double2 *sys, double2 *res, int NMq, int reps, int timed, long long *cyc, double *sink)
{
    extern __shared__ __align__(16) double2 sm[];
    __shared__ double2 colbuf[2 * 2 * T2_NRANK], rowbuf[4 * 2 * T2_NRANK];
    __shared__ int sing, stop;
    const int tid = threadIdx.x, lda = 2 * NMq;
    if (tid == 0) stop = 0;
    __syncthreads();
    if (tid >= 128) {
        // neighbours: independent DFMA chains until the solver group is done
        double a0 = tid * 1e-3, a1 = 1.0, a2 = 2.0, a3 = 3.0, a4 = 0.5, a5 = 0.7, a6 = 0.9, a7 = 1.1;
        volatile int *vs = &stop;
        while (!*vs) {
#pragma unroll
            for (int q = 0; q < 16; q++) {
                a0 = fma(a0, 1.0000001, 1e-9); a1 = fma(a1, 1.0000001, 1e-9); a2 = fma(a2, 1.0000001, 1e-9); a3 = fma(a3, 1.0000001, 1e-9);
                a4 = fma(a4, 1.0000001, 1e-9); a5 = fma(a5, 1.0000001, 1e-9); a6 = fma(a6, 1.0000001, 1e-9); a7 = fma(a7, 1.0000001, 1e-9);
            }
        }
        sink[blockIdx.x * blockDim.x + tid] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
        return;
    }
    const int cl = tid >> 6, warp = tid >> 5, lane = tid & 31;
    long long sec[4] = {0, 0, 0, 0};
    long long t0 = 0, tot = 0;
    for (int r = 0; r < reps; r++) {
        for (int e = tid; e < 2 * NMq * lda; e += 128) sm[e] = sys[e];
        asm volatile("bar.sync 3, 128;" ::: "memory");
        t0 = clock64();
        double2 *aug = sm + cl * NMq * lda;
        if (timed) solve_rows_timed(aug, lda, aug + NMq, lda, aug + NMq, lda, NMq, warp & 1, lane, 1 + cl, colbuf + cl * 2 * T2_NRANK,
                                    rowbuf + warp * 2 * T2_NRANK, &sing, sec);
        else t2_solve_rows(aug, lda, aug + NMq, lda, aug + NMq, lda, NMq, warp & 1, lane, 1 + cl, colbuf + cl * 2 * T2_NRANK,
                           rowbuf + warp * 2 * T2_NRANK, &sing);
        asm volatile("bar.sync 3, 128;" ::: "memory");
        tot += clock64() - t0;
    }
    if (tid == 0) {
        stop = 1;
        cyc[blockIdx.x * 8] = tot;
        for (int q = 0; q < 4; q++) cyc[blockIdx.x * 8 + 1 + q] = sec[q];
    }
    for (int e = tid; e < 2 * NMq * lda; e += 128) res[e] = sm[e];
}

int main(int argc, char **argv)
{
    const int NMq = argc > 1 ? atoi(argv[1]) : 17, reps = 50;
    const int lda = 2 * NMq, n = 2 * NMq * lda;
    std::vector<double2> h(n);
    srand(1);
    for (int c = 0; c < 2; c++)
        for (int r = 0; r < NMq; r++)
            for (int j = 0; j < lda; j++) {
                double2 v = make_double2(rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5);
                if (j == r) v.x += 3.0;
                h[(c * NMq + r) * lda + j] = v;
            }
    double2 *d_sys, *d_res;
    long long *d_cyc;
    double *d_sink;
    cudaMalloc(&d_sys, n * sizeof(double2)); cudaMalloc(&d_res, n * sizeof(double2));
    cudaMalloc(&d_cyc, 148 * 8 * sizeof(long long)); cudaMalloc(&d_sink, 148 * 1024 * sizeof(double));
    cudaMemcpy(d_sys, h.data(), n * sizeof(double2), cudaMemcpyHostToDevice);
    for (int nbusy : {0, 4, 8}) {
        for (int timed = 0; timed < 2; timed++) {
            k_bench<<<148, 128 + 32 * nbusy, n * sizeof(double2)>>>(d_sys, d_res, NMq, reps, timed, d_cyc, d_sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long hc[8];
            cudaMemcpy(hc, d_cyc, sizeof(hc), cudaMemcpyDeviceToHost);
            printf("NMq %d, %d FP64-busy neighbour warps, %s: %.0f cycles per step", NMq, nbusy, timed ? "timed copy" : "product code",
                   hc[0] / (double)reps / NMq);
            if (timed) printf("  [col publish + barrier %.0f, pivot search %.0f, pivot row %.0f, update %.0f]", hc[1] / (double)reps / NMq,
                              hc[2] / (double)reps / NMq, hc[3] / (double)reps / NMq, hc[4] / (double)reps / NMq);
            printf("\n");
        }
    }
    // residual check of the last run: L X = R
    std::vector<double2> out(n);
    cudaMemcpy(out.data(), d_res, n * sizeof(double2), cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int c = 0; c < 2; c++)
        for (int r = 0; r < NMq; r++)
            for (int q = 0; q < NMq; q++) {
                double sx = 0, sy = 0;
                for (int j = 0; j < NMq; j++) {
                    double2 a = h[(c * NMq + r) * lda + j], x = out[(c * NMq + j) * lda + NMq + q];
                    sx += a.x * x.x - a.y * x.y; sy += a.x * x.y + a.y * x.x;
                }
                double2 b = h[(c * NMq + r) * lda + NMq + q];
                worst = fmax(worst, fmax(fabs(sx - b.x), fabs(sy - b.y)));
            }
    printf("residual max |L X - R| = %.2e\n", worst);
    return 0;
}
