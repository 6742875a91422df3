// hs_regsolve.cuh -- register-resident Gauss-Jordan for the parity-class systems [A | B] -> A^-1 B of the staged pipeline
// (T = -RgQ Q^-1, pytmatrix `TT`), written for throughput: many small systems in flight per SM.
//
// Layout: lane = (row r, row group q); a warp holds G = 32 / ROWS row groups, NW warps work on one system, so the system has
// NW * G column slots of CPL columns each: the left half of the slots holds A (n <= ROWS columns, zero padded), the right
// half B.  Each lane keeps its row's CPL columns of its slot in registers for the whole elimination; nothing but the pivot
// column (NW > 1: through 16 bytes of shared memory per row and one named barrier per step; NW = 1: shuffles) and the pivot
// row (shuffles inside the warp: the pivot row's entries of a slot live in another lane of the same warp) is exchanged.
// Same pivot rule as the reference-order solver (solve_multi, hs_tmat.cuh): largest |re| + |im| among the rows not used
// yet, ties to the lowest position; rows are never moved nor scaled (vp = position after the reference's swaps, sc = the
// pending 1 / pivot: with W = sc V the true rows, V_r -= (V_r[k] / V_P[k]) V_P holds for every pending scale).
#pragma once
#include "hs_tmat.cuh"

namespace hs {

#ifndef HS_HOST_EMU

template <int ROWS, int CPL, int NW>
struct RegSolve {
    static constexpr int G = 32 / ROWS;       // row groups per warp
    static constexpr int NSLOT = NW * G;      // column slots of CPL columns
    static constexpr int NS = NSLOT * CPL / 2;  // columns of A (= columns of B) held
    static_assert(NS >= ROWS, "the slots must hold 2 * ROWS columns");
    static_assert(NSLOT % 2 == 0 || NSLOT == 1, "A and B halves must fall on slot boundaries");

    double2 v[CPL];
    double2 sc;
    int vp;

    // element (r, jg) of the augmented system in memory (jg < NS: A column jg, else B column jg - NS); zero outside n
    template <class LD>
    __device__ __forceinline__ void load(const SysDesc &d, int w, int lane, LD ld)
    {
        const int r = lane % ROWS, q = lane / ROWS, slot = w * G + q;
        const int n = d.n;
        const long rowoff = (long)(d.off + d.stride * r) * d.ld;
#pragma unroll
        for (int j = 0; j < CPL; j++) {
            const int jg = slot * CPL + j;
            const int col = jg < NS ? jg : jg - NS;
            double2 x = make_double2(0.0, 0.0);
            if (r < n && col < n) x = ld(d.base + rowoff + (jg < NS ? d.off + d.stride * col : d.rhs + d.off + d.stride * col));
            v[j] = x;
        }
        vp = r;
        sc = make_double2(1.0, 0.0);
    }

    // colbuf: NW > 1 only, [2][ROWS] double2 of shared memory per system; barid: named barrier of the system's NW warps
    __device__ __forceinline__ void eliminate(int n, int w, int lane, int barid, double2 *colbuf, int *sing)
    {
        const unsigned FULL = 0xffffffffu;
        const int r = lane % ROWS, q = lane / ROWS, slot = w * G + q;
        const bool realr = r < n;
        double2 f = v[0];  // this row's entry of the next pivot column, if this lane's slot owns it
        for (int k = 0; k < n; k++) {
            const int kslot = k / CPL;
            double2 c;
            if (NW > 1) {
                double2 *cb = colbuf + (k & 1) * ROWS;
                if (slot == kslot) cb[r] = f;
                asm volatile("bar.sync %0, %1;" ::"r"(barid), "r"(32 * NW) : "memory");
                c = cb[r];
            } else if (G > 1) {
                const int src = kslot * ROWS + r;
                c.x = __shfl_sync(FULL, f.x, src); c.y = __shfl_sync(FULL, f.y, src);
            } else {
                c = f;
            }
            const bool cand = realr && vp >= k;
            const unsigned long long bits = (unsigned long long)__double_as_longlong(fabs(c.x) + fabs(c.y));
            const unsigned hi = (unsigned)(bits >> 32) + 1u, lo = ((unsigned)bits & 0xFFFFFF00u) | (unsigned)(255 - vp);
            const unsigned m1 = __reduce_max_sync(FULL, cand ? hi : 0u);
            const bool top = cand && hi == m1;
            const unsigned m2 = __reduce_max_sync(FULL, top ? lo : 0u);
            const int P = (__ffs(__ballot_sync(FULL, top && lo == m2)) - 1) % ROWS;  // the row groups hold the same column
            if (m1 == 1u && (m2 >> 8) == 0u && w == 0 && lane == 0) *sing = 1;
            const double pvx = __shfl_sync(FULL, c.x, P), pvy = __shfl_sync(FULL, c.y, P);
            const int vpP = __shfl_sync(FULL, vp, P);
            const double rden = hs_rcp(pvx * pvx + pvy * pvy);
            const double2 inv = make_double2(pvx * rden, -pvy * rden);
            double2 cs = make_double2(-(c.x * inv.x - c.y * inv.y), -(c.x * inv.y + c.y * inv.x));
            if (r == P) { cs = make_double2(0.0, 0.0); sc = inv; }
            if (!realr) cs = make_double2(0.0, 0.0);
            double2 fn = f;
            const int jn = k + 1 - slot * CPL;  // local index of the next pivot column (if 0 <= jn < CPL)
            const int psrc = q * ROWS + P;      // lane of this warp that holds the pivot row's entries of this slot
            // columns that still matter: A columns beyond k, B columns below n (uniform per warp when G = 1)
            int jlo = 0, jhi = CPL;
            if (G == 1) {
                const int base = slot * CPL;
                if (base < NS) { jlo = k + 1 - base; jhi = n - base; }
                else jhi = n - (base - NS);
                jlo = jlo < 0 ? 0 : jlo;
                jhi = jhi > CPL ? CPL : jhi;
            } else {
                jhi = n < CPL ? n : CPL;  // the B slots need every column below n
            }
#pragma unroll
            for (int j = 0; j < CPL; j++) {
                if (j >= jlo && j < jhi) {
                    const double rx = __shfl_sync(FULL, v[j].x, psrc), ry = __shfl_sync(FULL, v[j].y, psrc);
                    double2 u = v[j];
                    u.x = fma(-cs.y, ry, fma(cs.x, rx, u.x));
                    u.y = fma(cs.y, rx, fma(cs.x, ry, u.y));
                    v[j] = u;
                    if (j == jn) fn = u;
                }
            }
            f = fn;
            if (vp == k) vp = vpP;
            if (r == P) vp = k;
        }
    }

    // solution entries held by this lane: calls st(row position vp, B column b, value) for the valid ones
    template <class ST>
    __device__ __forceinline__ void store(int n, int w, int lane, ST st) const
    {
        const int r = lane % ROWS, q = lane / ROWS, slot = w * G + q;
        if (r >= n) return;
#pragma unroll
        for (int j = 0; j < CPL; j++) {
            const int jg = slot * CPL + j;
            if (jg >= NS && jg - NS < n)
                st(vp, jg - NS, make_double2(v[j].x * sc.x - v[j].y * sc.y, v[j].x * sc.y + v[j].y * sc.x));
        }
    }
};

#endif  // HS_HOST_EMU

}  // namespace hs
