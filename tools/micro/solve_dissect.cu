// Where do the ~1 700 fixed cycles of one elimination step go?  solve_grp2 with parts switched off (timing only).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../hydroscatt_b200/csrc/hs_staged.cuh"
using namespace hs;
template <int DBG>
__device__ __forceinline__ void solve_dbg(const SysDesc d, int t, int nl, int barid, unsigned long long *keyslot, int *sing)
{
    const int n = d.n, nc = 2 * n;
    const int rstep = d.stride * d.ld;
    double2 *b0 = d.base + d.off * d.ld;
    const int ct = (t & 15) | ((t >> 5) << 4), rs = (t >> 4) & 1, ncl = nl >> 1;
    const int j0 = ct, j1 = ct + ncl;
    const int c0 = j0 < n ? d.off + d.stride * j0 : d.rhs + d.off + d.stride * (j0 - n);
    const int c1 = j1 < n ? d.off + d.stride * j1 : d.rhs + d.off + d.stride * (j1 - n);
    if (t == 0) keyslot[0] = 255ull | (1ull << 40);
    asm volatile("bar.sync %0, %1;" ::"r"(barid), "r"(nl) : "memory");
    for (int k = 0; k < n; k++) {
        const unsigned long long key = keyslot[k & 1];
        int p = 255 - (int)(key & 0xFFull);
        if (DBG & 8) p = k;
        if (DBG) p = p < 0 ? 0 : (p >= n ? n - 1 : p);
        const double2 *ck = b0 + (d.off + d.stride * k);
        const double2 pv = ck[p * rstep];
        const double den = pv.x * pv.x + pv.y * pv.y;
        double2 inv;
        if (DBG & 1) inv = make_double2(pv.x * 0.25, -pv.y * 0.25);
        else inv = make_double2(pv.x / den, -pv.y / den);
        const double2 fkk = ck[k * rstep];
        unsigned long long nkey = 0ull;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int j = h ? j1 : j0;
            const bool live = j > k && j < nc;
            double2 *cj = b0 + (h ? c1 : c0);
            double2 a = make_double2(0.0, 0.0), b = a;
            if (live) { a = cj[k * rstep]; b = cj[p * rstep]; }
            __syncwarp();
            if (live) {
                const double2 nk = make_double2(b.x * inv.x - b.y * inv.y, b.x * inv.y + b.y * inv.x);
                if (rs == 0) cj[k * rstep] = nk;
                const bool next_col = (j == k + 1) && (k + 1 < n);
                const double2 *fr = ck + rs * rstep;
                double2 *vr = cj + rs * rstep;
                const int rstep2 = 2 * rstep;
                if (!(DBG & 2)) {
#pragma unroll 4
                    for (int r = rs; r < n; r += 2, fr += rstep2, vr += rstep2) {
                        if (r == k) continue;
                        const double2 f = (r == p) ? fkk : *fr;
                        double2 v = (r == p) ? a : *vr;
                        v.x -= f.x * nk.x - f.y * nk.y;
                        v.y -= f.x * nk.y + f.y * nk.x;
                        *vr = v;
                        if (!(DBG & 8) && next_col && r > k) { const unsigned long long ky = pivot_key(v, r); nkey = ky > nkey ? ky : nkey; }
                    }
                }
            }
            if (j1 >= nc) break;
        }
        if (!(DBG & 8)) {
            const unsigned long long ok = __shfl_xor_sync(0xffffffffu, nkey, 16);
            nkey = ok > nkey ? ok : nkey;
        }
        if (rs == 0 && (j0 == k + 1 || j1 == k + 1)) keyslot[(k + 1) & 1] = (DBG & 8) ? (255ull - (k + 1)) | (1ull << 40) : nkey;
        if (DBG & 4) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(barid), "r"(nl) : "memory");
    }
}
template <int DBG>
__global__ void __launch_bounds__(256, 2) k(const double2 *src, int NM, int reps, long long *cyc)
{
    extern __shared__ __align__(16) double2 sys[];
    __shared__ unsigned long long keyslot[4][2];
    __shared__ SysDesc sd[2];
    __shared__ int sing;
    const int tid = threadIdx.x, warp = tid >> 5, ldc = 2 * NM;
    long long tot = 0;
    for (int rep = 0; rep < reps; rep++) {
        for (int t = tid; t < 4 * NM * NM; t += 256) sys[t] = src[t];
        if (tid < 2) { SysDesc d; d.base = sys + (long)tid * NM * ldc; d.n = NM; d.ld = ldc; d.off = 0; d.stride = 1; d.rhs = NM; sd[tid] = d; }
        __syncthreads();
        long long t0 = clock64();
        const int q = warp >> 2;
        solve_dbg<DBG>(sd[q], tid & 127, 128, 1 + q, keyslot[q], &sing);
        __syncthreads();
        tot += clock64() - t0;
    }
    if (tid == 0) cyc[blockIdx.x] = tot;
}

template <int DBG> double run(const double2 *d, int NM, long long *cyc)
{
    const int smem = 65536, reps = 20;
    cudaFuncSetAttribute(k<DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<DBG><<<148, 256, smem>>>(d, NM, reps, cyc);
    long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    return (double)h / reps / NM;
}
int main()
{
    for (int NM : {8, 16, 32}) {
        std::vector<double2> h(4 * NM * NM);
        srand(1);
        for (size_t i = 0; i < h.size(); i++) { h[i] = make_double2(rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5); if ((int)(i % (2 * NM)) == (int)((i / (2 * NM)) % NM)) h[i].x += 4.0; }
        double2 *d; long long *cyc;
        cudaMalloc(&d, h.size() * 16); cudaMemcpy(d, h.data(), h.size() * 16, cudaMemcpyHostToDevice); cudaMalloc(&cyc, 148 * 8);
        printf("NM %2d cycles/step: full %5.0f | no pivot search %5.0f | +no division %5.0f | +no row loop %5.0f | +no named barrier %5.0f | only no-division %5.0f | only no-barrier %5.0f  (%s)\n", NM,
               run<0>(d, NM, cyc), run<8>(d, NM, cyc), run<9>(d, NM, cyc), run<11>(d, NM, cyc), run<15>(d, NM, cyc), run<1>(d, NM, cyc), run<4>(d, NM, cyc), cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
