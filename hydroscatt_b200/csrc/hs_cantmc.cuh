// hs_cantmc.cuh -- Monte-Carlo angular moments of the canting-angle distribution on the device, consuming the SAME random stream
// as the reference: scattering.compute_angular_moments (hydroscatt/scattering.py:70-133) draws
//   rng = numpy.random.default_rng(seed);  alpha = rng.uniform(0, 360, n_part)            (one vector, reused for every sigma)
//   for sigma in canting_angle:  beta = rng.normal(scale=sigma, size=n_part)             (500 000 normals per sigma, up to 800 sigmas)
// i.e. numpy's PCG64 (128-bit LCG, XSL-RR output) followed by the 256-layer ziggurat of random_standard_normal, which consumes a
// DATA-DEPENDENT number of 64-bit draws per sample (1 in 98.8 % of the cases).  The stream is made parallel like this:
//   * PCG64 jumps ahead in O(log n) (LCG power), so any stream position can be entered directly;
//   * k_cmc_walk: one thread per block of CMC_B stream positions walks the samples that start in its block assuming the block is
//     entered at its first position, records which positions start a sample (bitmap), how many there are and by how many draws the
//     last one overruns into the next block;
//   * the true entry of block b is the overrun of block b - 1 (an orbit entered a few positions late merges with the recorded one at
//     the next sample start they share -- checked, never assumed); k_cmc_fix re-walks the few blocks entered late up to the merge
//     and corrects their sample count; a prefix sum gives every block the index of its first sample;
//   * k_cmc_emit: one thread per block walks its samples again and accumulates the six moments of its sigma(s); k_cmc_reduce sums
//     the per-block partial sums of a sigma in a fixed order (deterministic result).
// Tables: hs_ziggurat.h (generated from the installed numpy by tools/gen_ziggurat.py, whose Python restatement is checked draw for
// draw against numpy).  Differences to numpy that remain: libm vs CUDA sin / cos / exp / log1p (<= 2 ulp) and the summation order.
#pragma once
#include "hs_compat.cuh"
#include "hs_ziggurat.h"

namespace hs {

#ifndef HS_HOST_EMU

#define CMC_B 1024                 // stream positions per block (one thread)
#define CMC_WORDS (CMC_B / 32)

struct U128 { unsigned long long lo, hi; };
__host__ __device__ __forceinline__ U128 u128(unsigned long long hi, unsigned long long lo) { U128 r; r.lo = lo; r.hi = hi; return r; }
__device__ __forceinline__ U128 mul128(U128 a, U128 b)
{
    U128 r;
    r.lo = a.lo * b.lo;
    r.hi = __umul64hi(a.lo, b.lo) + a.lo * b.hi + a.hi * b.lo;
    return r;
}
__device__ __forceinline__ U128 add128(U128 a, U128 b)
{
    U128 r;
    r.lo = a.lo + b.lo;
    r.hi = a.hi + b.hi + (r.lo < a.lo ? 1ull : 0ull);
    return r;
}
#define PCG_MULT_HI 0x2360ED051FC65DA4ULL
#define PCG_MULT_LO 0x4385DF649FCCF645ULL

struct Pcg { U128 s, inc; };
// numpy's next64: step, then XSL-RR of the NEW state
__device__ __forceinline__ unsigned long long pcg_next(Pcg &g)
{
    g.s = add128(mul128(g.s, u128(PCG_MULT_HI, PCG_MULT_LO)), g.inc);
    const unsigned long long x = g.s.hi ^ g.s.lo;
    const unsigned r = (unsigned)(g.s.hi >> 58);
    return (x >> r) | (x << ((64u - r) & 63u));
}
__device__ __forceinline__ double pcg_double(Pcg &g) { return (double)(pcg_next(g) >> 11) * (1.0 / 9007199254740992.0); }
// state after `delta` steps (pcg_advance_lcg_128)
__device__ inline void pcg_advance(Pcg &g, unsigned long long delta)
{
    U128 cur_mult = u128(PCG_MULT_HI, PCG_MULT_LO), cur_plus = g.inc, acc_mult = u128(0, 1), acc_plus = u128(0, 0);
    while (delta) {
        if (delta & 1ull) { acc_mult = mul128(acc_mult, cur_mult); acc_plus = add128(mul128(acc_plus, cur_mult), cur_plus); }
        cur_plus = mul128(add128(cur_mult, u128(0, 1)), cur_plus);
        cur_mult = mul128(cur_mult, cur_mult);
        delta >>= 1;
    }
    g.s = add128(mul128(acc_mult, g.s), acc_plus);
}

// random_standard_normal (numpy/random/src/distributions/distributions.c); *ndraws += the 64-bit draws it consumed
__device__ inline double zig_normal(Pcg &g, int *ndraws)
{
    for (;;) {
        unsigned long long r = pcg_next(g);
        (*ndraws)++;
        const int idx = (int)(r & 0xff);
        r >>= 8;
        const int sign = (int)(r & 1);
        const unsigned long long rabs = (r >> 1) & 0x000fffffffffffffULL;
        double x = (double)rabs * __longlong_as_double((long long)ZIG_WI_BITS[idx]);
        if (sign) x = -x;
        if (rabs < ZIG_KI[idx]) return x;
        if (idx == 0) {
            for (;;) {
                const double xx = -HS_ZIG_INV_R * log1p(-pcg_double(g));
                const double yy = -log1p(-pcg_double(g));
                (*ndraws) += 2;
                if (yy + yy > xx * xx) return ((rabs >> 8) & 1) ? -(HS_ZIG_R + xx) : HS_ZIG_R + xx;
            }
        } else {
            const double f0 = __longlong_as_double((long long)ZIG_FI_BITS[idx - 1]), f1 = __longlong_as_double((long long)ZIG_FI_BITS[idx]);
            (*ndraws)++;
            if ((f0 - f1) * pcg_double(g) + f1 < exp(-0.5 * x * x)) return x;
        }
    }
}

struct CmcArgs {
    Pcg g0;                       // generator state of default_rng(seed) before any draw
    int n_part, n_cant;
    long long nblocks;
    const double *sigma;          // [n_cant] canting angles (deg)
    double sin_t0, cos_t0;        // thet0 = pi / 2 - ele pi / 180
    double *ca, *sa;              // [n_part] cos(alpha), sin(alpha)
    unsigned *bitmap;             // [nblocks][CMC_WORDS] sample starts of the orbit entered at the first position
    int *count0, *exit0, *entry, *cnt;   // [nblocks]
    long long *first;             // [nblocks] index of the block's first sample (exclusive prefix sum of cnt)
    double *partial;              // [nblocks][2][6] moment sums of the (at most two) sigmas a block touches
    int *flag;                    // != 0: an orbit did not merge inside a block (retry with another block size) / not enough blocks
};

// alpha = uniform(0, 360) * pi / 180: stream positions 0 .. n_part - 1, one draw each
__global__ void k_cmc_alpha(CmcArgs a)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n_part) return;
    Pcg g = a.g0;
    pcg_advance(g, (unsigned long long)t);
    const double u = pcg_double(g);
    const double al = (0.0 + 360.0 * u) * 3.141592653589793 / 180.0;
    double s, c;
    sincos(al, &s, &c);
    a.ca[t] = c;
    a.sa[t] = s;
}

__global__ void k_cmc_walk(CmcArgs a)
{
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.nblocks) return;
    Pcg g = a.g0;
    pcg_advance(g, (unsigned long long)a.n_part + (unsigned long long)b * CMC_B);
    unsigned words[CMC_WORDS];
#pragma unroll
    for (int w = 0; w < CMC_WORDS; w++) words[w] = 0u;
    int pos = 0, count = 0;
    while (pos < CMC_B) {
        // (register array with a dynamic index would go to local memory: set the bit through a predicated sweep instead)
        const int w = pos >> 5;
        const unsigned bit = 1u << (pos & 31);
#pragma unroll
        for (int q = 0; q < CMC_WORDS; q++) if (q == w) words[q] |= bit;
        int nd = 0;
        zig_normal(g, &nd);
        pos += nd;
        count++;
    }
#pragma unroll
    for (int w = 0; w < CMC_WORDS; w++) a.bitmap[b * CMC_WORDS + w] = words[w];
    a.count0[b] = count;
    a.exit0[b] = pos - CMC_B;
}

// blocks entered late: walk from the true entry until the orbit meets a recorded sample start, correct the sample count
__global__ void k_cmc_fix(CmcArgs a)
{
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.nblocks) return;
    const int e = b == 0 ? 0 : a.exit0[b - 1];
    a.entry[b] = e;
    if (e == 0) { a.cnt[b] = a.count0[b]; return; }
    const unsigned *bm = a.bitmap + b * CMC_WORDS;
    Pcg g = a.g0;
    pcg_advance(g, (unsigned long long)a.n_part + (unsigned long long)b * CMC_B + (unsigned long long)e);
    int pos = e, walked = 0;
    while (pos < CMC_B) {
        if ((bm[pos >> 5] >> (pos & 31)) & 1u) {
            int before = 0;  // recorded starts before pos
            for (int w = 0; w < (pos >> 5); w++) before += __popc(bm[w]);
            before += __popc(bm[pos >> 5] & ((1u << (pos & 31)) - 1u));
            a.cnt[b] = walked + a.count0[b] - before;
            return;
        }
        int nd = 0;
        zig_normal(g, &nd);
        pos += nd;
        walked++;
    }
    // left the block without meeting the recorded orbit: the overrun of THIS block may differ from exit0 -> give up (the caller
    // retries; with 1 024 positions per block and ~99 % of them sample starts this does not happen in practice)
    a.cnt[b] = walked;
    if (pos - CMC_B != a.exit0[b]) atomicExch(a.flag, 1);
}

// exclusive prefix sum of cnt over the blocks (one CTA; nblocks <= a few 10^5)
__global__ void k_cmc_scan(CmcArgs a)
{
    __shared__ long long part[1024];
    const int tid = threadIdx.x;
    const long long per = (a.nblocks + 1023) / 1024, lo = tid * per, hi = lo + per < a.nblocks ? lo + per : a.nblocks;
    long long s = 0;
    for (long long b = lo; b < hi; b++) s += a.cnt[b];
    part[tid] = s;
    __syncthreads();
    if (tid == 0) {
        long long acc = 0;
        for (int q = 0; q < 1024; q++) { const long long v = part[q]; part[q] = acc; acc += v; }
        if (acc < (long long)a.n_cant * a.n_part) atomicExch(a.flag, 2);  // not enough stream covered
    }
    __syncthreads();
    s = part[tid];
    for (long long b = lo; b < hi; b++) { a.first[b] = s; s += a.cnt[b]; }
}

__global__ void k_cmc_emit(CmcArgs a)
{
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.nblocks) return;
    double *out = a.partial + b * 12;
#pragma unroll
    for (int q = 0; q < 12; q++) out[q] = 0.0;
    const long long total = (long long)a.n_cant * a.n_part;
    long long j = a.first[b];
    if (j >= total) return;
    const int e = a.entry[b];
    Pcg g = a.g0;
    pcg_advance(g, (unsigned long long)a.n_part + (unsigned long long)b * CMC_B + (unsigned long long)e);
    int i = (int)(j / a.n_part), t = (int)(j - (long long)i * a.n_part);
    const int i_first = i;
    double sig = a.sigma[i];
    double s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s7 = 0;
    int pos = e;
    while (pos < CMC_B && j < total) {
        int nd = 0;
        const double z = zig_normal(g, &nd);
        pos += nd;
        const double beta = (0.0 + sig * z) * 3.141592653589793 / 180.0;
        double sb, cb;
        sincos(beta, &sb, &cb);
        const double sa = a.sa[t];
        const double x = cb * a.sin_t0 - sb * a.cos_t0 * a.ca[t];
        const double a1 = x * x, a2 = sb * sb * sa * sa;
        s1 += a1; s2 += a2; s3 += a1 * a1; s4 += a2 * a2; s5 += a1 * a2; s7 += a1 - a2;
        j++;
        if (++t == a.n_part) {
            // next sigma (a block touches at most two: CMC_B <= n_part is required by the host)
            double *o = out + (i - i_first) * 6;
            o[0] = s1; o[1] = s2; o[2] = s3; o[3] = s4; o[4] = s5; o[5] = s7;
            s1 = s2 = s3 = s4 = s5 = s7 = 0.0;
            t = 0; i++;
            if (i < a.n_cant) sig = a.sigma[i];
        }
    }
    if (i - i_first < 2) {
        double *o = out + (i - i_first) * 6;
        o[0] = s1; o[1] = s2; o[2] = s3; o[3] = s4; o[4] = s5; o[5] = s7;
    }
}

// mean over the n_part samples of sigma i: its samples are [i n_part, (i + 1) n_part), held by a contiguous range of blocks
__global__ void k_cmc_reduce(CmcArgs a, double *out)
{
    const int i = blockIdx.x, lane = threadIdx.x;   // one warp per sigma
    const long long j0 = (long long)i * a.n_part, j1 = j0 + a.n_part;
    // first block whose sample range reaches j0 (binary search on `first`), then forward while the block starts before j1
    long long lo = 0, hi = a.nblocks - 1;
    while (lo < hi) { const long long mid = (lo + hi + 1) >> 1; if (a.first[mid] <= j0) lo = mid; else hi = mid - 1; }
    double s[6] = {0, 0, 0, 0, 0, 0};
    for (long long b = lo + lane; b < a.nblocks && a.first[b] < j1; b += 32) {
        const long long fb = a.first[b];
        if (fb + a.cnt[b] <= j0) continue;
        const int i_first = (int)(fb / a.n_part);
        const int slot = i - i_first;
        if (slot < 0 || slot > 1) continue;
        const double *p = a.partial + b * 12 + slot * 6;
#pragma unroll
        for (int q = 0; q < 6; q++) s[q] += p[q];
    }
#pragma unroll
    for (int q = 0; q < 6; q++) {
        for (int o = 16; o; o >>= 1) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
        if (lane == 0) out[i * 6 + q] = s[q] / (double)a.n_part;
    }
}

#endif  // HS_HOST_EMU

}  // namespace hs
