// hs_tilesolve.cuh -- register-tiled Gauss-Jordan for the parity-class systems [A | B] -> A^-1 B that the assembly kernels
// leave in shared memory (T = -RgQ Q^-1, pytmatrix `TT`).
//
// A lane holds an RT x CT tile of the augmented system in registers: rows rb + 8 i (i < RT, rb = lane & 7: cyclic, so the
// shrinking set of unused rows stays spread over the lanes), columns CT cb + j (cb = thread >> 3).  One elimination step
// then exchanges RT multipliers and CT pivot-row entries per lane for RT x CT complex updates -- a row-per-lane or
// column-per-lane layout exchanges one value per update, and the exchange (shuffles or shared-memory round trips), not
// the arithmetic, is what bounds those (tools/micro/regsolve_bench.cu).  The exchange goes through three small arrays in
// shared memory (pivot column, pivot row, control words) and two named barriers per step among the system's NW warps.
// Pivot rule and bookkeeping as in the reference-order solver (solve_multi, hs_tmat.cuh): largest |re| + |im| among the
// rows not used yet, ties to the lowest position; rows are never moved nor scaled (vp = position after the reference's
// swaps; the pending scale 1 / pivot of the row that ends at position k is kept in invbuf[k] and applied on write-back).
#pragma once
#include "hs_tmat.cuh"

namespace hs {

#ifndef HS_HOST_EMU

// shared-memory scratch of one system
template <int ROWS, int NCOLS>
struct TileSolveBuf {
    double2 col[2][ROWS]; // pivot column of the step (raw entries), ping-pong: the update of step k reads it while step k + 1 is published
    double2 row[NCOLS];   // pivot row of the step (raw entries)
    double2 inv[ROWS];    // 1 / pivot of step k
    int ctl[4];           // pivot row, its position
};

template <int ROWS, int RT, int CT, int NW>
struct TileSolve {
    static constexpr int RB = ROWS / RT;             // row blocks = lanes per column block
    static constexpr int NCB = 32 * NW / RB;         // column blocks
    static constexpr int NCOLS = NCB * CT;           // augmented columns held
    static constexpr int NS = NCOLS / 2;             // columns of A (= of B)
    static_assert(RB == 8, "a column block is held by an aligned octet of lanes");
    static_assert(NS >= ROWS, "the tiles must hold 2 * ROWS columns");
    typedef TileSolveBuf<ROWS, NCOLS> Buf;

    // t = thread index inside the system's NW warps; the system lives in shared memory (SysDesc addressing)
    static __device__ __forceinline__ void solve(const SysDesc d, int t, int barid, Buf *buf, int *sing)
    {
        const int n = d.n;
        const int rb = t & 7, cb = t >> 3;
        const unsigned octmask = 0xFFu << (t & 24);
        const unsigned b0 = (unsigned)__cvta_generic_to_shared(d.base);
        // the live column blocks are packed into the first octets: A blocks 0 .. nsb - 1, then the B blocks, so that whole
        // warps hold no live column when n is small (and skip the update) and the A warps drop out as their columns pass k
        const int nsb = (n + CT - 1) / CT;
        const bool isA = cb < nsb, dead = cb >= 2 * nsb;
        const int cbl = isA ? cb : cb - nsb;  // block index inside A or B
        auto addr = [&](int r, int col, bool a_part) -> unsigned {  // byte address of element (r, col) of A or of B
            const int cc = a_part ? d.off + d.stride * col : d.rhs + d.off + d.stride * col;
            return b0 + 16u * (unsigned)((d.off + d.stride * r) * d.ld + cc);
        };
        auto bar = [&]() {
            if (NW > 1) asm volatile("bar.sync %0, %1;" ::"r"(barid), "r"(32 * NW) : "memory");
            else __syncwarp();
        };
        double2 v[RT][CT];
        int vp[RT];
#pragma unroll
        for (int i = 0; i < RT; i++) {
            const int r = rb + RB * i;
            vp[i] = r;
#pragma unroll
            for (int j = 0; j < CT; j++) {
                const int col = cbl * CT + j;
                const bool ok = r < n && col < n && !dead;
                v[i][j] = ok ? st_lds128(addr(r, col, isA)) : make_double2(0.0, 0.0);
            }
        }
        const unsigned colb0 = (unsigned)__cvta_generic_to_shared(&buf->col[0][0]), rowb = (unsigned)__cvta_generic_to_shared(buf->row);
        const int nkb = (n + CT - 1) / CT;
        for (int kb = 0; kb < nkb; kb++) {
#pragma unroll
            for (int jj = 0; jj < CT; jj++) {
                const int k = kb * CT + jj;
                if (k >= n) break;
                const unsigned colb = colb0 + 16u * (unsigned)((k & 1) * ROWS);
                // ---- owner octet of column k: publish the column, find the pivot ----
                if (cb == kb) {
                    unsigned long long key = 0ull;
                    int wi = 0;
#pragma unroll
                    for (int i = 0; i < RT; i++) {
                        const int r = rb + RB * i;
                        if (RB * i < n) {
                            const double2 c = v[i][jj];
                            st_sts128(colb + 16u * (unsigned)r, c);
                            if (r < n && vp[i] >= k) {
                                // order-preserving key (pivot_key): never 0, the low byte holds 255 - position
                                const unsigned long long ky = ((unsigned long long)__double_as_longlong(fabs(c.x) + fabs(c.y)) & ~0xFFull) | (unsigned long long)(255 - vp[i]);
                                if (ky > key) { key = ky; wi = i; }
                            }
                        }
                    }
                    unsigned long long best = key;
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) {
                        const unsigned long long ot = __shfl_xor_sync(octmask, best, o);
                        best = ot > best ? ot : best;
                    }
                    if (key == best && key != 0ull) {  // positions are distinct, so exactly one (lane, slot) holds the maximum
                        int pr = rb, pvp = 0;
#pragma unroll
                        for (int i = 0; i < RT; i++) if (i == wi) { pr = rb + RB * i; pvp = vp[i]; }
                        buf->ctl[0] = pr; buf->ctl[1] = pvp;
                        if ((best >> 8) == 0ull) *sing = 1;
                    }
                }
                bar();
                const int P = buf->ctl[0], vpP = buf->ctl[1];
                const double2 pv = st_lds128(colb + 16u * (unsigned)P);
                const double rden = hs_rcp(pv.x * pv.x + pv.y * pv.y);
                const double2 inv = make_double2(pv.x * rden, -pv.y * rden);
                if (t == 0) buf->inv[k] = inv;
                // ---- holders of the pivot row publish their columns ----
                if (rb == (P & 7)) {
                    const int iP = P >> 3;
#pragma unroll
                    for (int i = 0; i < RT; i++)
                        if (i == iP) {
#pragma unroll
                            for (int j = 0; j < CT; j++) st_sts128(rowb + 16u * (unsigned)(cb * CT + j), v[i][j]);
                        }
                }
                bar();
                // ---- update: columns of A beyond k and the valid columns of B; rows other than the pivot row ----
                const int c0 = cb * CT;
                const bool live = isA ? (c0 + CT - 1 > k) : !dead;
                if (__any_sync(0xffffffffu, live)) {
                    double2 pr[CT];
#pragma unroll
                    for (int j = 0; j < CT; j++) pr[j] = st_lds128(rowb + 16u * (unsigned)(c0 + j));
#pragma unroll
                    for (int i = 0; i < RT; i++) {
                        if (RB * i < n) {
                            const int r = rb + RB * i;
                            const double2 c = st_lds128(colb + 16u * (unsigned)r);
                            double2 cs = make_double2(-(c.x * inv.x - c.y * inv.y), -(c.x * inv.y + c.y * inv.x));
                            if (r == P || r >= n) cs = make_double2(0.0, 0.0);
#pragma unroll
                            for (int j = 0; j < CT; j++) {
                                double2 u = v[i][j];
                                u.x = fma(-cs.y, pr[j].y, fma(cs.x, pr[j].x, u.x));
                                u.y = fma(cs.y, pr[j].x, fma(cs.x, pr[j].y, u.y));
                                v[i][j] = u;
                            }
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < RT; i++) {
                    if (vp[i] == k) vp[i] = vpP;
                    if (rb + RB * i == P) vp[i] = k;
                }
            }
        }
        // every warp is past its last read of the system before solution rows overwrite right-hand sides
        bar();
        const unsigned invb = (unsigned)__cvta_generic_to_shared(buf->inv);
#pragma unroll
        for (int i = 0; i < RT; i++) {
            const int r = rb + RB * i;
            if (r < n) {
                const double2 sc = st_lds128(invb + 16u * (unsigned)vp[i]);
#pragma unroll
                for (int j = 0; j < CT; j++) {
                    const int col = cbl * CT + j;
                    if (!isA && !dead && col < n)
                        st_sts128(addr(vp[i], col, false), make_double2(v[i][j].x * sc.x - v[i][j].y * sc.y, v[i][j].x * sc.y + v[i][j].y * sc.x));
                }
            }
        }
    }
};

#endif  // HS_HOST_EMU

}  // namespace hs
