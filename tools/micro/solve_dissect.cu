// Where do the ~1 700 fixed cycles of one elimination step go?  solve_grp2 with parts switched off (timing only).
#include <cstdio>
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
                if (rs == 0) { cj[k * rstep] = nk; if (p != k) cj[p * rstep] = a; }
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

// candidate: rows in batches of 4 with all loads issued before the first store (the compiler cannot hoist a load above the
// previous iteration's store: possible aliasing), pivot candidates tracked as (truncated magnitude, row) instead of a packed key
__device__ __forceinline__ void solve_grp3(const SysDesc d, int t, int nl, int barid, unsigned long long *keyslot, int *sing)
{
    const int n = d.n, nc = 2 * n;
    const int rstep = d.stride * d.ld;
    double2 *b0 = d.base + d.off * d.ld;
    const int ct = (t & 15) | ((t >> 5) << 4), rs = (t >> 4) & 1, ncl = nl >> 1;
    const int j0 = ct, j1 = ct + ncl;
    const int c0 = j0 < n ? d.off + d.stride * j0 : d.rhs + d.off + d.stride * (j0 - n);
    const int c1 = j1 < n ? d.off + d.stride * j1 : d.rhs + d.off + d.stride * (j1 - n);
    if (t < 32) {
        unsigned long long key = 0ull;
        const double2 *cc = b0 + d.off;
        for (int r = t; r < n; r += 32) { const unsigned long long ky = pivot_key(cc[r * rstep], r); key = ky > key ? ky : key; }
        for (int o = 16; o; o >>= 1) { const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, o); key = ok > key ? ok : key; }
        if (t == 0) keyslot[0] = key;
    }
    asm volatile("bar.sync %0, %1;" ::"r"(barid), "r"(nl) : "memory");
    const int rstep2 = 2 * rstep;
    for (int k = 0; k < n; k++) {
        const unsigned long long key = keyslot[k & 1];
        const int p = 255 - (int)(key & 0xFFull);
        if ((key >> 8) == 0ull && t == 0) *sing = 1;
        const double2 *ck = b0 + (d.off + d.stride * k);
        const double2 pv = ck[p * rstep];
        const double den = pv.x * pv.x + pv.y * pv.y;
        const double2 inv = make_double2(pv.x / den, -pv.y / den);
        const double2 fkk = ck[k * rstep];
        double best = -1.0;  // truncated |re| + |im| of the best candidate of column k + 1 among this lane's rows
        int bi = 0;
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
                if (rs == 0) { cj[k * rstep] = nk; if (p != k) cj[p * rstep] = a; }
                const bool next_col = (j == k + 1) && (k + 1 < n);
                for (int r0 = rs; r0 < n; r0 += 8) {
                    double2 f[4], v[4];
                    bool ok[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int r = r0 + 2 * q;
                        ok[q] = r < n && r != k;
                        const int rr = r < n ? r : rs;
                        f[q] = ck[rr * rstep];
                        v[q] = cj[rr * rstep];
                        if (r == p) { f[q] = fkk; v[q] = a; }
                    }
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        v[q].x -= f[q].x * nk.x - f[q].y * nk.y;
                        v[q].y -= f[q].x * nk.y + f[q].y * nk.x;
                    }
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int r = r0 + 2 * q;
                        if (ok[q]) {
                            cj[r * rstep] = v[q];
                            if (next_col && r > k) {
                                // same order as pivot_key: magnitude with the low 8 mantissa bits cleared, smaller row wins ties
                                const double av = __longlong_as_double(__double_as_longlong(fabs(v[q].x) + fabs(v[q].y)) & ~0xFFll);
                                if (av > best) { best = av; bi = r; }
                            }
                        }
                    }
                }
            }
            if (j1 >= nc) break;
        }
        unsigned long long nkey = best >= 0.0 ? ((unsigned long long)__double_as_longlong(best) | (unsigned long long)(255 - bi)) : 0ull;
        const unsigned long long okk = __shfl_xor_sync(0xffffffffu, nkey, 16);
        nkey = okk > nkey ? okk : nkey;
        if (rs == 0 && (j0 == k + 1 || j1 == k + 1)) keyslot[(k + 1) & 1] = nkey;
        asm volatile("bar.sync %0, %1;" ::"r"(barid), "r"(nl) : "memory");
    }
    (void)rstep2;
}
template <int V>
__global__ void __launch_bounds__(256, 2) kv(const double2 *src, int NM, int reps, long long *cyc, double2 *dump)
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
        if (V == 2) solve_grp2(sd[q], tid & 127, 128, 1 + q, keyslot[q], &sing);
        else solve_grp3(sd[q], tid & 127, 128, 1 + q, keyslot[q], &sing);
        __syncthreads();
        tot += clock64() - t0;
    }
    if (tid == 0) cyc[blockIdx.x] = tot;
    if (blockIdx.x == 0) for (int t = tid; t < 4 * NM * NM; t += 256) dump[t] = sys[t];
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
        {
            double2 *dm[2]; cudaMalloc(&dm[0], h.size() * 16); cudaMalloc(&dm[1], h.size() * 16);
            cudaFuncSetAttribute(kv<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
            cudaFuncSetAttribute(kv<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
            long long c2 = 0, c3 = 0;
            kv<2><<<296, 256, 65536>>>(d, NM, 20, cyc, dm[0]); cudaMemcpy(&c2, cyc, 8, cudaMemcpyDeviceToHost);
            kv<3><<<296, 256, 65536>>>(d, NM, 20, cyc, dm[1]); cudaMemcpy(&c3, cyc, 8, cudaMemcpyDeviceToHost);
            std::vector<double2> r0(h.size()), r1(h.size());
            cudaMemcpy(r0.data(), dm[0], h.size() * 16, cudaMemcpyDeviceToHost); cudaMemcpy(r1.data(), dm[1], h.size() * 16, cudaMemcpyDeviceToHost);
            double md = 0; for (size_t i = 0; i < h.size(); i++) { if ((int)(i % (2 * NM)) < NM) continue; md = fmax(md, fmax(fabs(r0[i].x - r1[i].x), fabs(r0[i].y - r1[i].y))); }
            printf("      2 CTA/SM: solve_grp2 %5.0f  solve_grp3 %5.0f cycles/step, max diff %.1e (%s)\n", (double)c2 / 20 / NM, (double)c3 / 20 / NM, md, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
