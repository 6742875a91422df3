// Microbenchmark of the in-CTA Gauss-Jordan variants of hs_staged.cuh on synthetic systems (cycles per elimination step).
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../../hydroscatt_b200/csrc/hs_staged.cuh"
using namespace hs;

// mode 3: solve_grp2; mode 0: solve_grp (m>=1 layout: 2 systems x 128 lanes); 1: solve_sub (warp per system); 2: solve_multi (lock-step, 256 thr)
template <int MODE>
__global__ void __launch_bounds__(256, 2) k(const double2 *src, int NM, int reps, long long *cyc, double2 *dump)
{
    extern __shared__ __align__(16) double2 sys[];
    __shared__ unsigned long long keyslot[4][2];
    __shared__ SysDesc sd[HS_MAXSYS];
    __shared__ SolveCtl sc;
    __shared__ int sing;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, ldc = 2 * NM;
    long long tot = 0;
    for (int rep = 0; rep < reps; rep++) {
        for (int t = tid; t < 4 * NM * NM; t += 256) sys[t] = src[t];
        if (tid < 2) { SysDesc d; d.base = sys + (long)tid * NM * ldc; d.n = NM; d.ld = ldc; d.off = 0; d.stride = 1; d.rhs = NM; sd[tid] = d; }
        __syncthreads();
        long long t0 = clock64();
        if (MODE == 0) {
            const int q = warp >> 2;
            solve_grp(sd[q], tid & 127, 128, 1 + q, keyslot[q], &sing);
        } else if (MODE == 3) {
            const int q = warp >> 2;
            solve_grp2s(sd[q], tid & 127, 128, 1 + q, keyslot[q], &sing);
        } else if (MODE == 1) {
            if (warp < 2) solve_sub(sd[warp], lane, 32, NM, &sing);
        } else {
            Grp<false> grp; grp.tid = tid; grp.nt = 256;
            solve_multi(grp, sd, 2, &sc, &sing);
        }
        __syncthreads();
        tot += clock64() - t0;
    }
    if (tid == 0) cyc[blockIdx.x] = tot;
    if (dump && blockIdx.x == 0) for (int t = tid; t < 4 * NM * NM; t += 256) dump[t] = sys[t];
}

int main()
{
    for (int NM : {8, 14, 24, 32}) {
        std::vector<double2> h(4 * NM * NM);
        srand(1);
        for (int c = 0; c < 2; c++)
            for (int r = 0; r < NM; r++)
                for (int j = 0; j < 2 * NM; j++) {
                    double2 v = make_double2(rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5);
                    if (j == r) v.x += 4.0;
                    h[(c * NM + r) * 2 * NM + j] = v;
                }
        double2 *d; long long *cyc;
        cudaMalloc(&d, h.size() * 16); cudaMemcpy(d, h.data(), h.size() * 16, cudaMemcpyHostToDevice);
        cudaMalloc(&cyc, 148 * 4 * 8);
        const int smem = 65536, reps = 20;
        for (int per = 1; per <= 2; per++) {
            long long hc[3] = {0, 0, 0};
            cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            double2 *dmp[2]; cudaMalloc(&dmp[0], h.size() * 16); cudaMalloc(&dmp[1], h.size() * 16);
            cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            long long h3 = 0;
            k<0><<<148 * per, 256, smem>>>(d, NM, reps, cyc, nullptr); cudaMemcpy(&hc[0], cyc, 8, cudaMemcpyDeviceToHost);
            k<1><<<148 * per, 256, smem>>>(d, NM, reps, cyc, nullptr); cudaMemcpy(&hc[1], cyc, 8, cudaMemcpyDeviceToHost);
            k<2><<<148 * per, 256, smem>>>(d, NM, reps, cyc, dmp[0]); cudaMemcpy(&hc[2], cyc, 8, cudaMemcpyDeviceToHost);
            k<3><<<148 * per, 256, smem>>>(d, NM, reps, cyc, dmp[1]); cudaMemcpy(&h3, cyc, 8, cudaMemcpyDeviceToHost);
            std::vector<double2> r0(h.size()), r1(h.size());
            cudaMemcpy(r0.data(), dmp[0], h.size() * 16, cudaMemcpyDeviceToHost); cudaMemcpy(r1.data(), dmp[1], h.size() * 16, cudaMemcpyDeviceToHost);
            double md = 0; for (size_t i = 0; i < h.size(); i++) { int col = i % (2 * NM); if (col < NM) continue; md = fmax(md, fmax(fabs(r0[i].x - r1[i].x), fabs(r0[i].y - r1[i].y))); }
            printf("NM %2d, %d CTA/SM: cycles/step  solve_grp %6.0f  solve_grp2s %6.0f (max diff vs solve_multi %.1e)  solve_sub %6.0f  solve_multi %6.0f   (%s)\n", NM, per,
                   (double)hc[0] / reps / NM, (double)h3 / reps / NM, md, (double)hc[1] / reps / NM, (double)hc[2] / reps / NM, cudaGetErrorString(cudaGetLastError()));
            cudaFree(dmp[0]); cudaFree(dmp[1]);
        }
        cudaFree(d); cudaFree(cyc);
    }
    return 0;
}
