// hs_staged.cuh -- single-layer EBCM T-matrix as a staged pipeline: one kernel per stage, work items =
// (particle, trial configuration (NMAX, NGAUSS), azimuthal mode).
//
// Same algorithm as hs_tmat.cuh (pytmatrix `calctmat` = Mishchenko CONST/VARY/BESS/VIG/TMATR0/TMATR/TT, SURVEY.md
// rows P2-P7), different decomposition.  The fused kernel keeps one particle on one CTA for its whole life, so a
// W-band drop (NMAX 30: ~15 convergence trials, then 30 modes) is a 3.4 ms serial chain on one SM.  Here every trial
// and every mode is its own work item:
//   S  k_st_radial    thread = (item, quadrature node, function kind): j_n(kr), y_n(kr), j_n(m kr) + derivative forms
//   W  k_st_wigner    CTA = (item, mode), thread = node: Wigner d^n_{0m}, d/dtheta
//   A  k_st_assemble  CTA = (item, mode, 32x32 block of (n, n')): Q / RgQ integrals as FP64 DMMA (mma.sync.m8n8k4.f64)
//                     contractions over the quadrature nodes
//   G  k_st_solve     CTA = (item, mode): lock-step Gauss-Jordan of the parity-class systems, T block -> arena,
//                     Qsca/Qext of the m = 0 block -> host
// The host (hydrotm.cu: calctmat_staged) runs the NMAX / NGAUSS convergence controller between rounds and speculates
// trials, which are independent computations: only their Qsca/Qext feed the controller.
// Intermediates (radial tables, Wigner tables, [Q | RgQ] systems) travel through L2-resident global scratch.
#pragma once
#include "hs_tmat.cuh"

namespace hs {

struct StItem {
    int p, nmax, g, gp;        // particle, trial order, Gauss nodes on the half range, padded node count (multiple of 4)
    int nmodes, first_mi;      // 1 (trial: m = 0 only) or nmax + 1 (final: every mode); first mode-item index
    long long tab_off;         // doubles: 8 node arrays [gp] then 8 radial tables [nmax][gp]
    long long wig_off;         // doubles: per mode d1 [NM][gp], d2 [NM][gp]
    long long sys_off;         // complex: per mode [2][NM][2 NM]
    long long t_off;           // complex offset of the packed T in the caller's arena (set on the device), -1 = none
};
struct StModeItem { int item, m; };          // travels packed in one int: (m << 24) | item
__host__ __device__ inline int st_pack_mode(int item, int m) { return (m << 24) | item; }
__host__ __device__ inline StModeItem st_unpack_mode(int pm) { StModeItem r; r.item = pm & 0xFFFFFF; r.m = pm >> 24; return r; }
struct StRadTask { int item, kn; };        // kn = node0 << 2 | kind (0 = j + node arrays, 1 = y, 2 = j of the complex argument)
struct StATask { int mi; short sr, sc; };  // packed (item, mode), 32-wide row range, 32-wide column range of k = n - nmin
struct StResult { double qsca, qext; long long t_off; int sing, pad; };

struct StArgs {
    const StItem *items;
    const int *modes;        // every (item, mode) of the round, packed (k_st_wigner's work list)
    const double *axi, *lam, *mr, *mi, *eps;
    double rat;
    const double *gx, *gw;   // Gauss table (see TmArgs)
    const double *gfx, *gfw; // cylinders: full n-point rules, n = 1.., rule n at offset n (n - 1) / 2, nodes ascending
    int np;                  // shape: -1 spheroid, -2 finite circular cylinder (eps = diameter / length)
    const double *cm, *dd;   // cm[m] = prod_{i<=m} sqrt((2i-1)/2i); dd[n] = sqrt((2n+1)/(n(n+1)))
    double *tab, *wig;
    double2 *sys;
    double *tarena;
    long long arena_cap;
    unsigned long long *arena_top;
    long long *toff;
    StResult *res;
    int fuse;                // solve modes with NM <= 32 inside k_st_assemble (default) or in k_st_solve
#define ST_FUSE_RESERVED 16  // bit of `fuse`: result records cleared and arena slots reserved by the planner (device-driven rounds)
#ifndef ST_LEGACY_FUSED_SOLVERS
#define ST_LEGACY_FUSED_SOLVERS 0  // 1: keep the round-1 solvers (solve_grp2s, solve_rows16) selectable for the fused solves (fuse bits 2 / 4 / 8)
#endif
#ifndef ST_WIDE0_CTAS
#define ST_WIDE0_CTAS 2            // resident 256-thread CTAs per SM of the wide m = 0 assembly kernel (3 = 80 registers, 196 B of spills: measured equal, 5.01 vs 4.95 ms per W2 call)
#endif
#define ST_FUSE_TILE 32      // bit of `fuse`: register-tiled Gauss-Jordan (hs_tilesolve.cuh) for the fused solves of k_st_assemble
    unsigned long long *prof;  // HYDROTM_PROF=1: phase cycle counters of k_st_assemble [prologue, chunks, epilogue, solve, T write, #CTAs, sum NM, sum g]
};

enum { NA_X = 0, NA_S, NA_SS, NA_DR, NA_DDR, NA_DRR, NA_DRI, NA_RR, NA_COUNT };

__host__ __device__ inline long long st_tab_size(int nmax, int gp) { return (long long)(NA_COUNT + 8 * nmax) * gp; }
__host__ __device__ inline long long st_modes_nm_prefix(int nmax, int m)  // sum of NM over modes < m
{
    if (m == 0) return 0;
    return (long long)nmax + (long long)(m - 1) * (nmax + 1) - (long long)(m - 1) * m / 2;
}
__host__ __device__ inline long long st_sumsq(long long n) { return n * (n + 1) * (2 * n + 1) / 6; }
__host__ __device__ inline long long st_modes_nm2_prefix(int nmax, int m)  // sum of NM^2 over modes < m
{
    if (m == 0) return 0;
    return (long long)nmax * nmax + st_sumsq(nmax) - st_sumsq(nmax - m + 1);
}
__host__ __device__ inline long long st_wig_mode_off(int nmax, int m, int gp) { return 2LL * gp * st_modes_nm_prefix(nmax, m); }
__host__ __device__ inline long long st_sys_mode_off(int nmax, int m) { return 4LL * st_modes_nm2_prefix(nmax, m); }
__host__ __device__ inline long long st_t_mode_off(int nmax, int m) { return 2LL * st_modes_nm2_prefix(nmax, m); }

#ifndef HS_HOST_EMU

// Reciprocal for the dependent chains (downward ratio recurrences, Legendre recurrence, pivot inverse): the hardware seed
// (upper 32 bits) and two Newton steps, ~60 cycles of latency and 5 instructions against ~135 cycles / ~30 instructions of
// the IEEE division; 1-2 ulp.  Arguments are normal numbers here (a zero pivot is flagged singular before it is used).
__device__ __forceinline__ double hs_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}

}  // namespace hs
#include "hs_regsolve.cuh"
namespace hs {

// equal-volume radius of the particle (calctmat_particle: RAT != 1 -> SAREA)
__device__ inline double st_arev(double axi, double rat, double eps, int np = -1)
{
    if (fabs(rat - 1.0) > 1e-8) {
        if (np == -2) {  // SAREAC
            rat = pow(1.5 / eps, 1.0 / 3.0);
            rat = rat / sqrt((eps + 2.0) / (2.0 * eps));
        } else {
            double e, r;
            if (eps < 1.0) {
                e = sqrt(1.0 - eps * eps);
                r = 0.5 * (pow(eps, 2.0 / 3.0) + pow(eps, -1.0 / 3.0) * asin(e) / e);
            } else {
                e = sqrt(1.0 - 1.0 / (eps * eps));
                r = 0.25 * (2.0 * pow(eps, 2.0 / 3.0) + pow(eps, -4.0 / 3.0) * log((1.0 + e) / (1.0 - e)) / e);
            }
            rat = 1.0 / sqrt(r);
        }
    }
    return rat * axi;
}

// Quadrature node i (0 = next to the pole) of a trial with g nodes on the half range: x = cos(theta) < 0 and its weight.
// Spheroids: Gauss-Legendre on [-1, 0] (table of the 2g-point rule).  Cylinders (CONST, NP = -2): two Gauss segments
// split at the edge angle, [-1, xx] with g / 2 nodes and [xx, 0] with the rest, xx = -cos(atan(eps)).
__device__ __forceinline__ void st_node(const StArgs &a, int g, int i, double eps, double &x, double &w)
{
    if (a.np != -2) {
        const long goff = (long)g * (g - 1) / 2;
        x = -a.gx[goff + g - 1 - i];
        w = a.gw[goff + g - 1 - i];
        return;
    }
    const int ng1 = g / 2, ng2 = g - ng1;
    const double xx = -cos(atan(eps));
    if (i < ng1) {
        const long o = (long)ng1 * (ng1 - 1) / 2 + i;
        w = 0.5 * (xx + 1.0) * a.gfw[o];
        x = 0.5 * (xx + 1.0) * a.gfx[o] + 0.5 * (xx - 1.0);
    } else {
        const long o = (long)ng2 * (ng2 - 1) / 2 + (i - ng1);
        w = -0.5 * xx * a.gfw[o];
        x = -0.5 * xx * a.gfx[o] + 0.5 * xx;
    }
}

// surface at x = cos(theta) < 0: r^2 and (dr/dtheta) / r  (VARY: RSP1 spheroid, RSP3 cylinder); rev = equal-volume radius.
// The particle constants (one pow) are prepared once per thread, the surface is evaluated at several nodes.
struct StSurf {
    int np;
    double aa, ee, ee1;  // spheroid: a^2, eps^2, eps^2 - 1
    double h, ac;        // cylinder: half length, radius
    __device__ __forceinline__ void init(int np_, double rev, double eps)
    {
        np = np_;
        if (np != -2) {
            const double A = rev * pow(eps, 1.0 / 3.0);
            aa = A * A; ee = eps * eps; ee1 = ee - 1.0; h = 0.0; ac = 0.0;
        } else {
            h = rev * pow(2.0 / (3.0 * eps * eps), 1.0 / 3.0); ac = h * eps; aa = 0.0; ee = 0.0; ee1 = 0.0;
        }
    }
    __device__ __forceinline__ void at(double x, double &r2, double &drr) const
    {
        if (np != -2) {
            const double cc = x * x, ssq = 1.0 - cc, sth = sqrt(ssq);
            const double rq = 1.0 / (ssq + ee * cc);
            r2 = aa * rq;
            drr = rq * x * sth * ee1;
            return;
        }
        const double co = -x, si = sqrt(1.0 - co * co);
        double rad, rthet;
        if (si / co > ac / h) { rad = ac / si; rthet = -ac * co / (si * si); }
        else { rad = h / co; rthet = h * si / (co * co); }
        r2 = rad * rad;
        drr = -rthet / rad;
    }
};

// ---- S: radial functions.  Same recurrences as rjb / ryb / cjb, with the ratios parked in a thread-local array and the
// results written n-major ([n][gp], coalesced over the nodes of a warp).
__device__ inline void st_rjb(double x, double *y, double *u, int nmax, int nnmax, int st)
{
    double zz[HS_NPN1];
    const int l = nmax + nnmax;
    const double xx = 1.0 / x;
    double zc = 1.0 / ((double)(2 * l + 1) * xx);
    for (int i1 = l - 1; i1 >= 1; i1--) {
        zc = hs_rcp((double)(2 * i1 + 1) * xx - zc);
        if (i1 <= nmax) zz[i1 - 1] = zc;
    }
    const double z1 = zz[0];
    const double z0 = 1.0 / (xx - z1);
    const double y0 = z0 * cos(x) * xx;
    double yp = y0 * z1;
    u[0] = y0 - yp * xx;
    y[0] = yp;
#pragma unroll 4
    for (int i = 2; i <= nmax; i++) {
        const double yi = yp * zz[i - 1];
        u[(long)(i - 1) * st] = yp - (double)i * yi * xx;
        y[(long)(i - 1) * st] = yi;
        yp = yi;
    }
}

__device__ inline void st_ryb(double x, double *y, double *v, int nmax, int st)
{
    double s, c;
    sincos(x, &s, &c);
    const double x1 = 1.0 / x, x2 = x1 * x1, x3 = x2 * x1;
    const double y1 = -c * x2 - s * x1;
    y[0] = y1;
    v[0] = -x1 * (c + y1);
    if (nmax < 2) return;
    double ym2 = y1, ym1 = (-3.0 * x3 + x1) * c - 3.0 * x2 * s;  // y_1, y_2
    y[st] = ym1;
    v[st] = ym2 - 2.0 * x1 * ym1;
    for (int i = 2; i <= nmax - 1; i++) {
        const double yi = (double)(2 * i + 1) * x1 * ym1 - ym2;  // y_{i+1}
        y[(long)i * st] = yi;
        v[(long)i * st] = ym1 - (double)(i + 1) * x1 * yi;
        ym2 = ym1; ym1 = yi;
    }
}

__device__ inline void st_cjb(double xr, double xi, double *yr, double *yi, double *ur, double *ui, int nmax, int nnmax,
                              int st)
{
    double zr_[HS_NPN1], zi_[HS_NPN1];
    const int l = nmax + nnmax;
    const double xrxi = 1.0 / (xr * xr + xi * xi);
    const double cxxr = xr * xrxi, cxxi = -xi * xrxi;
    double qf = 1.0 / (double)(2 * l + 1);
    double czr = xr * qf, czi = xi * qf;
    for (int i1 = l - 1; i1 >= 1; i1--) {
        qf = (double)(2 * i1 + 1);
        const double ar = qf * cxxr - czr;
        const double ai = qf * cxxi - czi;
        const double ari = hs_rcp(ar * ar + ai * ai);
        czr = ar * ari;
        czi = -ai * ari;
        if (i1 <= nmax) { zr_[i1 - 1] = czr; zi_[i1 - 1] = czi; }
    }
    const double z1r = zr_[0], z1i = zi_[0];
    double ar = cxxr - z1r, ai = cxxi - z1i;
    const double ari = 1.0 / (ar * ar + ai * ai);
    const double cz0r = ar * ari, cz0i = -ai * ari;
    double sxr, cxr;
    sincos(xr, &sxr, &cxr);
    const double cr = cxr * cosh(xi), ci = -sxr * sinh(xi);
    ar = cz0r * cr - cz0i * ci;
    ai = cz0i * cr + cz0r * ci;
    const double cy0r = ar * cxxr - ai * cxxi, cy0i = ai * cxxr + ar * cxxi;
    double p1r = cy0r * z1r - cy0i * z1i, p1i = cy0i * z1r + cy0r * z1i;
    ar = p1r * cxxr - p1i * cxxi;
    ai = p1i * cxxr + p1r * cxxi;
    yr[0] = p1r; yi[0] = p1i;
    ur[0] = cy0r - ar; ui[0] = cy0i - ai;
#pragma unroll 4
    for (int i = 2; i <= nmax; i++) {
        const double qi = (double)i;
        const double zr = zr_[i - 1], zi = zi_[i - 1];
        const double cyir = p1r * zr - p1i * zi;
        const double cyii = p1i * zr + p1r * zi;
        ar = cyir * cxxr - cyii * cxxi;
        ai = cyii * cxxr + cyir * cxxi;
        const long o = (long)(i - 1) * st;
        yr[o] = cyir; yi[o] = cyii;
        ur[o] = p1r - qi * ar; ui[o] = p1i - qi * ai;
        p1r = cyir; p1i = cyii;
    }
}

// one radial task (16 nodes of one function kind of one item) on a half-warp; l16 = lane within the half-warp
__device__ __forceinline__ void st_radial_task(const StArgs &a, const StRadTask t, int l16)
{
    const StItem it = a.items[t.item];
    const int kind = t.kn & 3, i = (t.kn >> 2) + l16;
    const int g = it.g, nmax = it.nmax, gp = it.gp;
    if (t.kn == 2 && l16 == 0 && !(a.fuse & ST_FUSE_RESERVED)) {
        // first task of the item: clear its result record and, for a final item, reserve the packed T in the caller's arena
        // (same bump allocator as the fused kernel; a retried final trial keeps its slot)
        long long o = -1;
        if (it.nmodes > 1) {
            o = it.t_off;
            if (o < 0) {
                const unsigned long long sz = (unsigned long long)tmat_size(nmax);
                const unsigned long long off = atomicAdd(a.arena_top, sz);
                o = (off + sz > (unsigned long long)a.arena_cap) ? -2 : (long long)off;
            }
        }
        const_cast<StItem *>(a.items)[t.item].t_off = o;
        a.res[t.item].t_off = o;
        a.res[t.item].sing = 0;
    }
    if (i >= g) return;
    const double PI = 3.14159265358979323846;
    const int p = it.p;
    const double lam = a.lam[p], mrr = a.mr[p], mri = a.mi[p], eps = a.eps[p];
    const double arev = st_arev(a.axi[p], a.rat, eps, a.np);
    const double kw = PI * 2.0 / lam;
    double x, wv, r, drr;
    StSurf surf;
    surf.init(a.np, arev, eps);
    st_node(a, g, i, eps, x, wv);
    surf.at(x, r, drr);
    const double xp = -x;
    const double v = sqrt(r) * kw;
    double *tab = a.tab + it.tab_off;
    double *tb = tab + (long)NA_COUNT * gp;
    const long tsz = (long)nmax * gp;
    // extra orders of the downward recurrences (VARY): from the largest k r over the nodes (spheroid: one of the end
    // nodes; cylinder: one of the two nodes next to the edge)
    double ta;
    {
        const int ia = a.np == -2 ? g / 2 - 1 : 0, ib = a.np == -2 ? g / 2 : g - 1;
        double xa, xb, wa, ra, rb, da;
        st_node(a, g, ia, eps, xa, wa); surf.at(xa, ra, da);
        st_node(a, g, ib, eps, xb, wa); surf.at(xb, rb, da);
        ta = fmax(sqrt(ra) * kw, sqrt(rb) * kw);
    }
    if (kind == 0) {
        const double v0 = 1.0 / (mrr * mrr + mri * mri);
        const double prr = mrr * v0, pri = -mri * v0;
        const double yv = 1.0 / (1.0 - x * x);
        const double vi = 1.0 / v;
        tab[NA_X * gp + i] = xp;
        tab[NA_S * gp + i] = sqrt(yv);
        tab[NA_SS * gp + i] = yv;
        tab[NA_DR * gp + i] = drr;
        tab[NA_DDR * gp + i] = vi;
        tab[NA_DRR * gp + i] = prr * vi;
        tab[NA_DRI * gp + i] = pri * vi;
        tab[NA_RR * gp + i] = wv * r;
        const double tmx = fmax(ta, (double)nmax);
        const int nnmax1 = (int)(1.2 * sqrt(tmx) + 3.0);
        st_rjb(v, tb + 0 * tsz + i, tb + 1 * tsz + i, nmax, nnmax1, gp);
    } else if (kind == 1) {
        st_ryb(v, tb + 2 * tsz + i, tb + 3 * tsz + i, nmax, gp);
    } else {
        double tbv = ta * sqrt(mrr * mrr + mri * mri);
        tbv = fmax(tbv, (double)nmax);
        int nnmax2 = (int)(tbv + 4.0 * pow(tbv, 0.33333) + 1.2 * sqrt(tbv));
        nnmax2 = nnmax2 - nmax + 5;
        st_cjb(v * mrr, v * mri, tb + 4 * tsz + i, tb + 5 * tsz + i, tb + 6 * tsz + i, tb + 7 * tsz + i, nmax, nnmax2, gp);
    }
}

__global__ void __launch_bounds__(128) k_st_radial(StArgs a, const StRadTask *tasks, int ntasks)
{
    // one half-warp per task of 16 nodes (most trial items have 10..16 nodes); the host orders the tasks kind-major, so
    // the two halves of a warp run the same recurrence on items of similar size
    const int wt = blockIdx.x * 8 + (threadIdx.x >> 4);
    if (wt >= ntasks) return;
    st_radial_task(a, tasks[wt], threadIdx.x & 15);
}

// Device-driven rounds (hs_staged_dev.h): the task lists and their lengths are produced on the device, so the stage
// kernels are launched with a fixed grid of resident CTAs that fetch tasks from a cursor until the list is empty.
struct StQueue { int count, cursor; };
__global__ void __launch_bounds__(128) k_st_radial_q(StArgs a, const StRadTask *tasks, StQueue *q)
{
    // the tasks are short and uniform (the planner orders them kind-major): a static stride over the resident warps costs
    // no atomic round trip per task pair (15 % of this kernel's stall samples with a cursor)
    const int lane = threadIdx.x & 31;
    const int n = q->count;
    const int nw = gridDim.x * (blockDim.x >> 5);
    for (int t0 = 2 * (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)); t0 < n; t0 += 2 * nw) {
        const int wt = t0 + (lane >> 4);
        if (wt < n) st_radial_task(a, tasks[wt], lane & 15);
        __syncwarp();
    }
}

// ---- W: Wigner functions of one (item, mode); rows k = n - nmin, n-major ----
// one half-warp per (item, mode): most modes have 10..16 nodes; larger ones loop over node chunks
__device__ __forceinline__ void st_wigner_task(const StArgs &a, int idx, bool live, double *sq, double *rsq, int l16)
{
    const StModeItem mi = st_unpack_mode(live ? a.modes[idx] : 0);
    const StItem it = a.items[mi.item];
    const int m = live ? mi.m : 0, nmax = it.nmax, g = live ? it.g : 0, gp = it.gp;
    const int nmin = m > 1 ? m : 1, NM = nmax - nmin + 1;
    double *d1 = a.wig + it.wig_off + st_wig_mode_off(nmax, m, gp);
    double *d2 = d1 + (long)NM * gp;
    const double eps = a.eps[it.p];
    if (live && m > 0) {
        const double qmm = (double)m * (double)m;
        for (int n = m + l16; n <= nmax + 1; n += 16) {
            const double v = sqrt((double)n * (double)n - qmm);
            sq[n] = v;
            rsq[n] = (n > m) ? 1.0 / v : 0.0;
        }
    }
    __syncwarp();
    for (int i = l16; i < g; i += 16) {
        double xn, wn;
        st_node(a, g, i, eps, xn, wn);
        const double x = -xn;
        const double qs = sqrt(1.0 - x * x), qs1 = 1.0 / qs;
        if (m == 0) {
            double a1 = 1.0, a2 = x;
            for (int n = 1; n <= nmax; n++) {
                const double qn = n, qn1 = n + 1, qn2 = 2 * n + 1;
                const double a3 = (qn2 * x * a2 - qn * a1) * hs_rcp(qn1);
                const double der = qs1 * (qn1 * qn * hs_rcp(qn2)) * (-a1 + a3);
                const double si = (n & 1) ? -1.0 : 1.0;
                d1[(long)(n - 1) * gp + i] = a2 * si;
                d2[(long)(n - 1) * gp + i] = -der * si;
                a1 = a2; a2 = a3;
            }
        } else {
            double qp = qs;
            for (int j = 2; j <= m; j++) qp = qp * qs;
            double a1 = 0.0, a2 = a.cm[m] * qp;
            for (int n = m; n <= nmax; n++) {
                const double qn = n, qn1 = n + 1, qn2 = 2 * n + 1;
                const double qnm = sq[n], qnm1 = sq[n + 1];
                const double a3 = (qn2 * x * a2 - qnm * a1) * rsq[n + 1];
                const double der = qs1 * (-qn1 * qnm * a1 + qn * qnm1 * a3) * hs_rcp(qn2);
                const double si = ((n + m) & 1) ? -1.0 : 1.0;
                d1[(long)(n - nmin) * gp + i] = a2 * si;
                d2[(long)(n - nmin) * gp + i] = -der * si;
                a1 = a2; a2 = a3;
            }
        }
    }
}

__global__ void __launch_bounds__(128) k_st_wigner(StArgs a, int n_mi)
{
    __shared__ double sq_[8][HS_NPN1 + 2], rsq_[8][HS_NPN1 + 2];
    const int hw = threadIdx.x >> 4;
    const int idx = blockIdx.x * 8 + hw;
    st_wigner_task(a, idx, idx < n_mi, sq_[hw], rsq_[hw], threadIdx.x & 15);
}

__global__ void __launch_bounds__(128) k_st_wigner_q(StArgs a, StQueue *q)
{
    __shared__ double sq_[8][HS_NPN1 + 2], rsq_[8][HS_NPN1 + 2];
    const int hw = threadIdx.x >> 4, lane = threadIdx.x & 31;
    const int n = q->count;
    const int nw = gridDim.x * (blockDim.x >> 5);
    for (int t0 = 2 * (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)); t0 < n; t0 += 2 * nw) {
        const int idx = t0 + (lane >> 4);
        st_wigner_task(a, idx, idx < n, sq_[hw], rsq_[hw], lane & 15);
        __syncwarp();
    }
}

__device__ __forceinline__ double2 st_lds128(unsigned addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void st_sts128(unsigned addr, double2 v)
{
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}

}  // namespace hs
#include "hs_tilesolve.cuh"
namespace hs {

// ---- Gauss-Jordan with row pivoting of one system by a (sub-)warp: lane = column of the augmented block ----
// Same pivot rule and arithmetic as solve_multi (hs_tmat.cuh); no CTA barrier: a lane only ever touches its own
// columns plus the (read-only during the step) pivot column, so one __syncwarp per elimination step is enough, and
// the pivot of the next step is found by the lane that has just updated that column.  `gw` lanes (8, 16 or 32) work
// on one system, so a warp solves 32 / gw small systems side by side; kmax = largest order among them.
__device__ __forceinline__ void solve_sub(const SysDesc d, int gl, int gw, int kmax, int *sing)
{
    const int n = d.n, nc = 2 * n;
    const long rstep = (long)d.stride * d.ld, row0 = (long)d.off * d.ld;
    auto colof = [&](int j) -> long { return (j < n ? d.off + d.stride * j : d.rhs + d.off + d.stride * (j - n)); };
    double2 *b0 = d.base + row0;
    unsigned long long key = 0ull;
    {
        const double2 *c0 = b0 + colof(0);
        for (int r = gl; r < n; r += gw) { const unsigned long long ky = pivot_key(c0[r * rstep], r); key = ky > key ? ky : key; }
        for (int o = gw >> 1; o; o >>= 1) { const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, o, gw); key = ok > key ? ok : key; }
    }
    for (int k = 0; k < kmax; k++) {
        unsigned long long nkey = 0ull;
        if (k < n) {
            const int p = 255 - (int)(key & 0xFFull);
            if ((key >> 8) == 0ull && gl == 0) *sing = 1;
            const double2 *ck = b0 + colof(k);
            const double2 pv = ck[p * rstep];
            const double den = pv.x * pv.x + pv.y * pv.y;
            const double2 inv = make_double2(pv.x / den, -pv.y / den);
            const double2 fkk = ck[k * rstep];
            for (int j = k + 1 + gl; j < nc; j += gw) {
                double2 *cj = b0 + colof(j);
                const double2 a = cj[k * rstep], b = cj[p * rstep];
                const double2 nk = make_double2(b.x * inv.x - b.y * inv.y, b.x * inv.y + b.y * inv.x);
                cj[k * rstep] = nk;
                if (p != k) cj[p * rstep] = a;
                const bool next_col = (j == k + 1) && (k + 1 < n);
#pragma unroll 4
                for (int r = 0; r < n; r++) {
                    if (r == k) continue;
                    // after the row swap, row p holds the old row k whose column-k entry is still at (k, k)
                    const double2 f = (r == p) ? fkk : ck[r * rstep];
                    double2 v = cj[r * rstep];
                    v.x -= f.x * nk.x - f.y * nk.y;
                    v.y -= f.x * nk.y + f.y * nk.x;
                    cj[r * rstep] = v;
                    if (next_col && r > k) { const unsigned long long ky = pivot_key(v, r); nkey = ky > nkey ? ky : nkey; }
                }
            }
        }
        key = __shfl_sync(0xffffffffu, nkey, 0, gw);  // column k + 1 belongs to lane 0 of the group
        __syncwarp();
    }
}


// solve_sub for systems that live in shared memory
__device__ __forceinline__ void solve_subs(const SysDesc d, int gl, int gw, int kmax, int *sing)
{
    const int n = d.n, nc = 2 * n;
    const long rstep = (long)d.stride * d.ld, row0 = (long)d.off * d.ld;
    auto colof = [&](int j) -> long { return (j < n ? d.off + d.stride * j : d.rhs + d.off + d.stride * (j - n)); };
    const unsigned b0 = (unsigned)__cvta_generic_to_shared(d.base + row0);  // explicit LDS / STS (see solve_grp2s)
    unsigned long long key = 0ull;
    {
        const unsigned c0 = b0 + 16u * (unsigned)colof(0);
        for (int r = gl; r < n; r += gw) { const unsigned long long ky = pivot_key(st_lds128(c0 + 16u * (unsigned)(r * rstep)), r); key = ky > key ? ky : key; }
        for (int o = gw >> 1; o; o >>= 1) { const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, o, gw); key = ok > key ? ok : key; }
    }
    for (int k = 0; k < kmax; k++) {
        unsigned long long nkey = 0ull;
        if (k < n) {
            const int p = 255 - (int)(key & 0xFFull);
            if ((key >> 8) == 0ull && gl == 0) *sing = 1;
            const unsigned ck = b0 + 16u * (unsigned)colof(k);
            const double2 pv = st_lds128(ck + 16u * (unsigned)(p * rstep));
            const double rden = hs_rcp(pv.x * pv.x + pv.y * pv.y);
            const double2 inv = make_double2(pv.x * rden, -pv.y * rden);
            const double2 fkk = st_lds128(ck + 16u * (unsigned)(k * rstep));
            for (int j = k + 1 + gl; j < nc; j += gw) {
                const unsigned cj = b0 + 16u * (unsigned)colof(j);
                const double2 a = st_lds128(cj + 16u * (unsigned)(k * rstep)), b = st_lds128(cj + 16u * (unsigned)(p * rstep));
                const double2 nk = make_double2(b.x * inv.x - b.y * inv.y, b.x * inv.y + b.y * inv.x);
                st_sts128(cj + 16u * (unsigned)(k * rstep), nk);
                if (p != k) st_sts128(cj + 16u * (unsigned)(p * rstep), a);
                const bool next_col = (j == k + 1) && (k + 1 < n);
#pragma unroll 4
                for (int r = 0; r < n; r++) {
                    if (r == k) continue;
                    // after the row swap, row p holds the old row k whose column-k entry is still at (k, k)
                    const double2 f = (r == p) ? fkk : st_lds128(ck + 16u * (unsigned)(r * rstep));
                    double2 v = st_lds128(cj + 16u * (unsigned)(r * rstep));
                    v.x -= f.x * nk.x - f.y * nk.y;
                    v.y -= f.x * nk.y + f.y * nk.x;
                    st_sts128(cj + 16u * (unsigned)(r * rstep), v);
                    if (next_col && r > k) { const unsigned long long ky = pivot_key(v, r); nkey = ky > nkey ? ky : nkey; }
                }
            }
        }
        key = __shfl_sync(0xffffffffu, nkey, 0, gw);  // column k + 1 belongs to lane 0 of the group
        __syncwarp();
    }
}



// The same elimination by a group of `nl` lanes (64 or 128 = 2 or 4 whole warps): two lanes of a warp share a column
// (even / odd rows), so a step is n / 2 row updates per lane; one named barrier per step; the two lanes that own
// column k + 1 publish the next pivot through shared memory.
__device__ __forceinline__ void solve_grp(const SysDesc d, int t, int nl, int barid, unsigned long long *keyslot, int *sing)
{
    const int n = d.n, nc = 2 * n;
    const long rstep = (long)d.stride * d.ld, row0 = (long)d.off * d.ld;
    auto colof = [&](int j) -> long { return (j < n ? d.off + d.stride * j : d.rhs + d.off + d.stride * (j - n)); };
    double2 *b0 = d.base + row0;
    // within a warp: lanes 0-15 take the even rows of 16 consecutive columns, lanes 16-31 the odd rows of the same columns
    // (each half-warp access is one contiguous 256-byte run: no bank conflicts whatever the row stride)
    const int ct = (t & 15) | ((t >> 5) << 4), rs = (t >> 4) & 1, ncl = nl >> 1;
    if (t < 32) {
        unsigned long long key = 0ull;
        const double2 *c0 = b0 + colof(0);
        for (int r = t; r < n; r += 32) { const unsigned long long ky = pivot_key(c0[r * rstep], r); key = ky > key ? ky : key; }
        for (int o = 16; o; o >>= 1) { const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, o); key = ok > key ? ok : key; }
        if (t == 0) keyslot[0] = key;
    }
    asm volatile("bar.sync %0, %1;" ::"r"(barid), "r"(nl) : "memory");
    for (int k = 0; k < n; k++) {
        const unsigned long long key = keyslot[k & 1];
        unsigned long long nkey = 0ull;
        const int p = 255 - (int)(key & 0xFFull);
        if ((key >> 8) == 0ull && t == 0) *sing = 1;
        const double2 *ck = b0 + colof(k);
        const double2 pv = ck[p * rstep];
        const double den = pv.x * pv.x + pv.y * pv.y;
        const double2 inv = make_double2(pv.x / den, -pv.y / den);
        const double2 fkk = ck[k * rstep];
        // all lanes of a warp run the same number of passes (the __syncwarp below is warp-wide)
        const int npass = (nc - k - 1 + ncl - 1) / ncl;
        for (int ps = 0; ps < npass; ps++) {
            const int j = k + 1 + ps * ncl + ct;
            const bool live = j < nc;
            double2 *cj = b0 + colof(live ? j : k);
            double2 a = make_double2(0.0, 0.0), b = a;
            if (live) { a = cj[k * rstep]; b = cj[p * rstep]; }
            __syncwarp();  // both lanes of the column have read rows k and p before one of them rewrites them
            if (!live) continue;
            const double2 nk = make_double2(b.x * inv.x - b.y * inv.y, b.x * inv.y + b.y * inv.x);
            if (rs == 0) cj[k * rstep] = nk;  // (row p is rewritten below by the lane that owns its parity: no second writer)
            const bool next_col = (j == k + 1) && (k + 1 < n);
#pragma unroll 4
            for (int r = rs; r < n; r += 2) {
                if (r == k) continue;
                // after the row swap, row p holds the old row k (value a; its column-k entry is still at (k, k))
                const double2 f = (r == p) ? fkk : ck[r * rstep];
                double2 v = (r == p) ? a : cj[r * rstep];
                v.x -= f.x * nk.x - f.y * nk.y;
                v.y -= f.x * nk.y + f.y * nk.x;
                cj[r * rstep] = v;
                if (next_col && r > k) { const unsigned long long ky = pivot_key(v, r); nkey = ky > nkey ? ky : nkey; }
            }
        }
        if (t < 32) {
            const unsigned long long ok = __shfl_xor_sync(0xffffffffu, nkey, 16);
            nkey = ok > nkey ? ok : nkey;
            if (t == 0) keyslot[(k + 1) & 1] = nkey;
        }
        asm volatile("bar.sync %0, %1;" ::"r"(barid), "r"(nl) : "memory");
    }
}


// Leaner variant of solve_grp (the elimination is instruction-bound, not latency-bound: ~1 900 cycles per step were
// measured for a step whose dependent latencies add up to ~450): fixed column ownership (lane pair ct owns columns ct,
// ct + nl/2: no per-step pointer arithmetic), 32-bit offsets, the pivot key is formed only by the pair that owns
// column k + 1.  Same pivot rule, same arithmetic, same results as solve_grp.
__device__ __forceinline__ void solve_grp2(const SysDesc d, int t, int nl, int barid, unsigned long long *keyslot, int *sing)
{
    const int n = d.n, nc = 2 * n;
    const int rstep = d.stride * d.ld;
    double2 *b0 = d.base + d.off * d.ld;
    const int ct = (t & 15) | ((t >> 5) << 4), rs = (t >> 4) & 1, ncl = nl >> 1;
    // this lane pair's columns (at most two: nc <= 2 * ncl)
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
    for (int k = 0; k < n; k++) {
        const unsigned long long key = keyslot[k & 1];
        const int p = 255 - (int)(key & 0xFFull);
        if ((key >> 8) == 0ull && t == 0) *sing = 1;
        const double2 *ck = b0 + (d.off + d.stride * k);
        const double2 pv = ck[p * rstep];
        const double den = pv.x * pv.x + pv.y * pv.y;
        const double2 inv = make_double2(pv.x / den, -pv.y / den);
        const double2 fkk = ck[k * rstep];
        unsigned long long nkey = 0ull;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int j = h ? j1 : j0;
            const bool live = j > k && j < nc;
            double2 *cj = b0 + (h ? c1 : c0);
            double2 a = make_double2(0.0, 0.0), b = a;
            if (live) { a = cj[k * rstep]; b = cj[p * rstep]; }
            __syncwarp();  // both lanes of the column have read rows k and p before one of them rewrites them
            if (live) {
                const double2 nk = make_double2(b.x * inv.x - b.y * inv.y, b.x * inv.y + b.y * inv.x);
                if (rs == 0) cj[k * rstep] = nk;  // (row p is rewritten below by the lane that owns its parity: no second writer)
                const bool next_col = (j == k + 1) && (k + 1 < n);
                const double2 *fr = ck + rs * rstep;
                double2 *vr = cj + rs * rstep;
                const int rstep2 = 2 * rstep;
#pragma unroll 4
                for (int r = rs; r < n; r += 2, fr += rstep2, vr += rstep2) {
                    if (r == k) continue;
                    // after the row swap, row p holds the old row k (value a; its column-k entry is still at (k, k))
                    const double2 f = (r == p) ? fkk : *fr;
                    double2 v = (r == p) ? a : *vr;
                    v.x -= f.x * nk.x - f.y * nk.y;
                    v.y -= f.x * nk.y + f.y * nk.x;
                    *vr = v;
                    if (next_col && r > k) { const unsigned long long ky = pivot_key(v, r); nkey = ky > nkey ? ky : nkey; }
                }
            }
            if (ncl >= nc) break;  // group-uniform: no lane has a second column
        }
        // the pair that owns column k + 1 publishes the next pivot
        const unsigned long long ok = __shfl_xor_sync(0xffffffffu, nkey, 16);
        nkey = ok > nkey ? ok : nkey;
        if (rs == 0 && (j0 == k + 1 || j1 == k + 1)) keyslot[(k + 1) & 1] = nkey;
        asm volatile("bar.sync %0, %1;" ::"r"(barid), "r"(nl) : "memory");
    }
}

// solve_grp2 for systems that live in shared memory
__device__ __forceinline__ void solve_grp2s(const SysDesc d, int t, int nl, int barid, unsigned long long *keyslot, int *sing)
{
    const int n = d.n, nc = 2 * n;
    const int rstep = d.stride * d.ld;
    // explicit shared-space accesses (32-bit addresses, LDS / STS): through the generic pointer of SysDesc the compiler emits
    // LD.E / ST.E, which resolve the address space at run time and cost ~3x the latency per row update
    const unsigned b0 = (unsigned)__cvta_generic_to_shared(d.base + d.off * d.ld);
    const int ct = (t & 15) | ((t >> 5) << 4), rs = (t >> 4) & 1, ncl = nl >> 1;
    // this lane pair's columns (at most two: nc <= 2 * ncl)
    const int j0 = ct, j1 = ct + ncl;
    const int c0 = j0 < n ? d.off + d.stride * j0 : d.rhs + d.off + d.stride * (j0 - n);
    const int c1 = j1 < n ? d.off + d.stride * j1 : d.rhs + d.off + d.stride * (j1 - n);
    if (t < 32) {
        unsigned long long key = 0ull;
        const unsigned cc = b0 + 16u * d.off;
        for (int r = t; r < n; r += 32) { const unsigned long long ky = pivot_key(st_lds128(cc + 16u * (r * rstep)), r); key = ky > key ? ky : key; }
        for (int o = 16; o; o >>= 1) { const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, o); key = ok > key ? ok : key; }
        if (t == 0) keyslot[0] = key;
    }
    asm volatile("bar.sync %0, %1;" ::"r"(barid), "r"(nl) : "memory");
    for (int k = 0; k < n; k++) {
        const unsigned long long key = keyslot[k & 1];
        const int p = 255 - (int)(key & 0xFFull);
        if ((key >> 8) == 0ull && t == 0) *sing = 1;
        const unsigned ck = b0 + 16u * (d.off + d.stride * k);
        const double2 pv = st_lds128(ck + 16u * (p * rstep));
        const double rden = hs_rcp(pv.x * pv.x + pv.y * pv.y);  // one reciprocal per step (the other variants divide twice: last-ulp differences)
        const double2 inv = make_double2(pv.x * rden, -pv.y * rden);
        const double2 fkk = st_lds128(ck + 16u * (k * rstep));
        unsigned long long nkey = 0ull;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int j = h ? j1 : j0;
            const bool live = j > k && j < nc;
            const unsigned cj = b0 + 16u * (h ? c1 : c0);
            double2 a = make_double2(0.0, 0.0), b = a;
            if (live) { a = st_lds128(cj + 16u * (k * rstep)); b = st_lds128(cj + 16u * (p * rstep)); }
            __syncwarp();  // both lanes of the column have read rows k and p before one of them rewrites them
            if (live) {
                const double2 nk = make_double2(b.x * inv.x - b.y * inv.y, b.x * inv.y + b.y * inv.x);
                if (rs == 0) st_sts128(cj + 16u * (k * rstep), nk);  // (row p is rewritten below by the lane that owns its parity: no second writer)
                const bool next_col = (j == k + 1) && (k + 1 < n);
                // rows in batches of 4: all loads of a batch are issued before its first store (an in-order thread otherwise
                // waits for each LDS before it can issue the next row's loads: ~100 cycles per row instead of ~35)
                const unsigned rstep2 = 32u * rstep;
                unsigned fr = ck + 16u * (rs * rstep), vr = cj + 16u * (rs * rstep);
                for (int r0 = rs; r0 < n; r0 += 8, fr += 4u * rstep2, vr += 4u * rstep2) {
                    double2 f[4], v[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int r = r0 + 2 * q;
                        const bool in = r < n;
                        f[q] = st_lds128(in ? fr + q * rstep2 : fr);
                        v[q] = st_lds128(in ? vr + q * rstep2 : vr);
                    }
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int r = r0 + 2 * q;
                        // after the row swap, row p holds the old row k (value a; its column-k entry is still at (k, k))
                        if (r == p) { f[q] = fkk; v[q] = a; }
                        v[q].x -= f[q].x * nk.x - f[q].y * nk.y;
                        v[q].y -= f[q].x * nk.y + f[q].y * nk.x;
                    }
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int r = r0 + 2 * q;
                        if (r < n && r != k) {
                            st_sts128(vr + q * rstep2, v[q]);
                            if (next_col && r > k) { const unsigned long long ky = pivot_key(v[q], r); nkey = ky > nkey ? ky : nkey; }
                        }
                    }
                }
            }
            if (ncl >= nc) break;  // group-uniform: no lane has a second column
        }
        // the pair that owns column k + 1 publishes the next pivot
        const unsigned long long ok = __shfl_xor_sync(0xffffffffu, nkey, 16);
        nkey = ok > nkey ? ok : nkey;
        if (rs == 0 && (j0 == k + 1 || j1 == k + 1)) keyslot[(k + 1) & 1] = nkey;
        asm volatile("bar.sync %0, %1;" ::"r"(barid), "r"(nl) : "memory");
    }
}


// ---- The same elimination out of REGISTERS (used by the fused solve of k_st_assemble; the two-layer kernel has a 17-row copy).
// Lane r keeps ROW r of the system (n <= 32); the 2 n columns of [A | B] are dealt in blocks of 16 to the W = 1, 2 or 4 warps of the
// group (2 n <= 16 W).  A step: the warp that owns pivot column k publishes it (one STS per lane, double-buffered, one named barrier
// of the group per step); every warp finds the pivot row with two warp reductions (same rule as solve_multi: largest |re| + |im|
// among the rows not used yet, low 8 mantissa bits dropped, ties to the lowest row); the pivot lane hands its 16 entries to its warp
// through shared memory and every lane updates its row:  V_r -= (V_r[k] / pivot) V_P.  Rows are neither moved nor scaled: `vp` is the
// position the row would have after the reference's row swaps and `sc` = 1 / pivot the pending scale of a used row (with W = sc V
// the true rows, the update of row r != P reads the same for every pending scale), so lane r ends up with solution row vp = sc V_r,
// which it writes back over the right-hand side.  No LDS -> DFMA -> STS round trip per row update, no arithmetic in one-lane
// branches: ~250 instructions per step and warp against ~2 000-3 000 cycles per step of the shared-memory variant.
__device__ __forceinline__ void solve_rows16(const SysDesc d, int w, int W, int lane, int barid, double2 *colbuf, double2 *rowbuf, int *sing)
{
    const unsigned FULL = 0xffffffffu;
    const int n = d.n, nc = 2 * n;
    const unsigned b0 = (unsigned)__cvta_generic_to_shared(d.base);
    // byte offset of element (r, j): row (off + stride r), column off + stride j (j < n) or rhs + off + stride (j - n)
    const unsigned rowoff = 16u * (unsigned)((d.off + d.stride * lane) * d.ld);
    double2 v[16];
    const bool realr = lane < n;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const int jg = 16 * w + j;
        double2 x = make_double2(0.0, 0.0);
        if (realr && jg < nc) x = st_lds128(b0 + rowoff + 16u * (unsigned)(jg < n ? d.off + d.stride * jg : d.rhs + d.off + d.stride * (jg - n)));
        v[j] = x;
    }
    int vp = lane;
    double2 f = v[0];                             // this row's entry of the next pivot column, if this warp owns it
    double2 sc = make_double2(1.0, 0.0);
    for (int k = 0; k < n; k++) {
        double2 *cb = colbuf + (k & 1) * 32;
        if (w == (k >> 4)) cb[lane] = f;
        asm volatile("bar.sync %0, %1;" ::"r"(barid), "r"(32 * W) : "memory");
        const double2 c = cb[lane];
        const bool cand = realr && vp >= k;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(fabs(c.x) + fabs(c.y));
        const unsigned hi = (unsigned)(bits >> 32) + 1u, lo = ((unsigned)bits & 0xFFFFFF00u) | (unsigned)(255 - vp);
        const unsigned m1 = __reduce_max_sync(FULL, cand ? hi : 0u);
        const bool top = cand && hi == m1;
        const unsigned m2 = __reduce_max_sync(FULL, top ? lo : 0u);
        const int P = __ffs(__ballot_sync(FULL, top && lo == m2)) - 1;
        if (m1 == 1u && (m2 >> 8) == 0u && w == 0 && lane == 0) *sing = 1;
        if (lane == P) {
#pragma unroll
            for (int j = 0; j < 16; j++) rowbuf[j] = v[j];
        }
        const double pvx = __shfl_sync(FULL, c.x, P), pvy = __shfl_sync(FULL, c.y, P);
        const int vpP = __shfl_sync(FULL, vp, P);
        const double rden = hs_rcp(pvx * pvx + pvy * pvy);
        const double2 inv = make_double2(pvx * rden, -pvy * rden);
        double2 cs = make_double2(-(c.x * inv.x - c.y * inv.y), -(c.x * inv.y + c.y * inv.x));
        if (lane == P) { cs = make_double2(0.0, 0.0); sc = inv; }
        if (!realr) cs = make_double2(0.0, 0.0);
        __syncwarp();
        double2 fn = f;
        const int jn = k + 1 - 16 * w;            // local index of the next pivot column (if 0 <= jn < 16)
#pragma unroll
        for (int j0 = 0; j0 < 16; j0 += 4) {
            double2 rw[4];
#pragma unroll
            for (int q = 0; q < 4; q++) rw[q] = rowbuf[j0 + q];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int j = j0 + q;
                double2 u = v[j];
                u.x = fma(-cs.y, rw[q].y, fma(cs.x, rw[q].x, u.x));
                u.y = fma(cs.y, rw[q].x, fma(cs.x, rw[q].y, u.y));
                v[j] = u;
                if (j == jn) fn = u;
            }
        }
        f = fn;
        if (vp == k) vp = vpP;
        if (lane == P) vp = k;
        __syncwarp();  // rowbuf is rewritten by the next step's pivot lane
    }
    // every warp is past its last read of the system before solution rows overwrite right-hand sides (row vp != row lane)
    asm volatile("bar.sync %0, %1;" ::"r"(barid), "r"(32 * W) : "memory");
    if (realr) {
        const unsigned outrow = 16u * (unsigned)((d.off + d.stride * vp) * d.ld);
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int jg = 16 * w + j;
            if (jg >= n && jg < nc)
                st_sts128(b0 + outrow + 16u * (unsigned)(d.rhs + d.off + d.stride * (jg - n)),
                          make_double2(v[j].x * sc.x - v[j].y * sc.y, v[j].x * sc.y + v[j].y * sc.x));
        }
    }
}

// ---- A: Q / RgQ integrals of one 32 x 32 block of (k1, k2) = (n1 - nmin, n2 - nmin) as DMMA contractions ----
// See assemble_tiled (hs_tmat.cuh) for the factorisation: X = sum over nodes of (left factor of (n1, node)) x (right
// factor of (n2, node)).  Left operand tables L_t[(k1, F)][node], F in {j, y}: P1 = d1 dF, P2 = d2 dF, P3 = d1 F,
// P4 = d2 F, P3a = n1(n1+1) P3.  Right operand tables R_t[(k2, re|im)][node]: q1..q5 for same-parity pairs, r1..r5 for
// opposite-parity pairs.  Rows are parity-sorted (even k: rows 0..31, odd k: rows 32..63) so that each parity block is
// a dense GEMM:  X12 = P1 q1 + P2 q2 + P3a q3,  X21 = P3 q4 + P4 q5,  X11 = P3 r1 + P4 r2,  X22 = P1 r3 + P2 r4 + P3a r5.
// A chunk of 8 nodes is staged in shared memory ([15 tables][64 rows][8 nodes], k4-group XOR-swizzled by row pair so
// the fragment loads are conflict-free); each warp owns one parity block's tile rows for both outputs (so the
// epilogue that combines X_A and X_B into Q and RgQ entries stays in registers).
#define ST_OP_DOUBLES (15 * 64 * 8)
#define ST_ASM_SMEM16 (15 * 32 * 8 * 8)  // KR = 16 variant: operand area 30 720 B (>= the 16 KB of its fused systems)
#define ST_ASM_SMEM 65536  // operand area (61 440 B), re-used for the [2][32][64] complex systems of a fused solve

__device__ __forceinline__ void st_dmma(double *c, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// KR = width of the (n, n') block handled by one CTA: 32 (256 threads, any NM, multi-block above 32) or 16 (128 threads,
// NM <= 16 only: half the registers and a quarter of the shared memory per CTA, so twice as many CTAs per SM for the
// majority of the modes).
template <bool M0, int KR>
__device__ __forceinline__ void st_assemble_task(const StArgs &a, const StATask tk, double *op)
{
    constexpr int THREADS = 8 * KR, TSTR = 16 * KR, TC = KR / 8, HMAX = KR / 2;  // table stride (doubles), tile columns, rows per parity
    long long tprof = a.prof ? clock64() : 0;
#define ST_PROF(slot) if (a.prof && threadIdx.x == 0) { const long long now_ = clock64(); atomicAdd(&a.prof[slot], (unsigned long long)(now_ - tprof)); tprof = now_; }
    const StModeItem mi = st_unpack_mode(tk.mi);
    const StItem it = a.items[mi.item];
    const int m = M0 ? 0 : mi.m;
    const int nmax = it.nmax, g = it.g, gp = it.gp;
    const int nmin = m > 1 ? m : 1, NM = nmax - nmin + 1;
    const int kr0 = KR * tk.sr, kc0 = KR * tk.sc;
    const double *nod = a.tab + it.tab_off;
    const double *tb = nod + (long)NA_COUNT * gp;
    const long tsz = (long)nmax * gp;
    const double *w1 = a.wig + it.wig_off + st_wig_mode_off(nmax, m, gp);
    const double *w2 = w1 + (long)NM * gp;
    const double qm = (double)m, qmm = qm * qm;
    const int tid = threadIdx.x, kk = tid >> 3, ii = tid & 7;
    const int warp = tid >> 5, lane = tid & 31;

    // warp -> parity block (prow, pcol) and tile rows [tr0, tr0 + NTR)
    constexpr int NTR = M0 ? 1 : 2;
    constexpr int TRS = (KR == 32 && !M0) ? 2 : 1;  // tile-row stride between a warp's accumulator slots
    int prow, pcol, tr0;
    // KR = 32: tile rows are dealt to a block's two warps round-robin (m >= 1: rows tr0, tr0 + 2), so half-empty blocks
    // stay balanced; KR = 16: one warp per block (m >= 1: both tile rows) or per (block, tile row) (m = 0)
    if (KR == 32) {
        if (M0) { prow = pcol = warp >> 2; tr0 = warp & 3; }
        else { const int beta = warp >> 1; prow = (beta == 1 || beta == 3) ? 1 : 0; pcol = (beta == 1 || beta == 2) ? 1 : 0; tr0 = warp & 1; }
    } else {
        if (M0) { prow = pcol = warp >> 1; tr0 = warp & 1; }
        else { prow = (warp == 1 || warp == 3) ? 1 : 0; pcol = (warp == 1 || warp == 2) ? 1 : 0; tr0 = 0; }
    }
    const bool even = prow == pcol;
    int hrow = (NM - kr0 - prow + 1) >> 1; hrow = hrow < 0 ? 0 : (hrow > HMAX ? HMAX : hrow);
    int hcol = (NM - kc0 - pcol + 1) >> 1; hcol = hcol < 0 ? 0 : (hcol > HMAX ? HMAX : hcol);
    const int TMv = (2 * hrow + 7) >> 3, TNv = (2 * hcol + 7) >> 3;

    double acc[NTR][TC][2][2];
#pragma unroll
    for (int r = 0; r < NTR; r++)
#pragma unroll
        for (int c = 0; c < TC; c++)
#pragma unroll
            for (int e = 0; e < 4; e++) acc[r][c][e >> 1][e & 1] = 0.0;

    // build-phase constants of this thread: rows (kk, F) / (kk, re|im) of every table, column ii (swizzled)
    const int rho = (kk & 1) * KR + 2 * (kk >> 1);
    const int col = ii ^ (((rho >> 1) & 1) << 2);
    // physical row = r ^ (r >= KR): 64-bit shared-memory accesses are served per half-warp, and a half-warp stores the row
    // pairs of one even and one odd k (kk = 2 h, 2 h + 1: rows r and KR + r), which fell into the same 8-word half of a bank
    // line; with the odd-parity block's row pairs swapped they alternate, and the fragment loads stay conflict-free
    const int rsw = (kk & 1);
    double *o0 = op + (rho ^ rsw) * 8 + col, *o1 = op + ((rho + 1) ^ rsw) * 8 + col;
    const int k1 = kr0 + kk, k2 = kc0 + kk;
    const int n1 = nmin + k1, n2 = nmin + k2;
    const double an1 = (double)(n1 * (n1 + 1)), an2 = (double)(n2 * (n2 + 1));
    const int fragrow = lane >> 2, fragk = lane & 3;

    ST_PROF(0)
    for (int i0 = 0; i0 < g; i0 += 8) {
        {
            const int i = i0 + ii;
            const bool vi = i < g;
            // left operand rows (k1, j | y).  Rows of k >= NM are never stored: they only feed accumulator rows / columns
            // that the epilogue discards (the kernel is shared-memory-bandwidth bound: mean NM is 14 of the 32 rows)
            if (k1 < NM) {
                double d1 = 0.0, d2 = 0.0, fj = 0.0, fdj = 0.0, fy = 0.0, fdy = 0.0;
                if (vi) {
                    const long o = (long)(n1 - 1) * gp + i;
                    d1 = w1[(long)k1 * gp + i]; d2 = w2[(long)k1 * gp + i];
                    fj = tb[o]; fdj = tb[tsz + o]; fy = tb[2 * tsz + o]; fdy = tb[3 * tsz + o];
                }
                o0[0 * TSTR] = d1 * fdj; o1[0 * TSTR] = d1 * fdy;  // P1
                o0[1 * TSTR] = d2 * fdj; o1[1 * TSTR] = d2 * fdy;  // P2
                const double p3j = d1 * fj, p3y = d1 * fy;
                o0[2 * TSTR] = p3j; o1[2 * TSTR] = p3y;            // P3
                o0[3 * TSTR] = d2 * fj; o1[3 * TSTR] = d2 * fy;    // P4
                o0[4 * TSTR] = an1 * p3j; o1[4 * TSTR] = an1 * p3y;  // P3a
            }
            // right operand rows (k2, re | im)
            if (k2 < NM) {
                double d1 = 0.0, d2 = 0.0, jr = 0.0, ji = 0.0, djr = 0.0, dji = 0.0;
                double rri = 0.0, uri = 0.0, ddri = 0.0, drri = 0.0, drii = 0.0, ssi = 0.0, si = 0.0;
                if (vi) {
                    const long o = (long)(n2 - 1) * gp + i;
                    d1 = w1[(long)k2 * gp + i]; d2 = w2[(long)k2 * gp + i];
                    jr = tb[4 * tsz + o]; ji = tb[5 * tsz + o]; djr = tb[6 * tsz + o]; dji = tb[7 * tsz + o];
                    rri = nod[NA_RR * gp + i]; uri = nod[NA_DR * gp + i]; ddri = nod[NA_DDR * gp + i];
                    drri = nod[NA_DRR * gp + i]; drii = nod[NA_DRI * gp + i];
                    if (!M0) { ssi = nod[NA_SS * gp + i]; si = nod[NA_S * gp + i]; }
                }
                const double jzr = jr * drri - ji * drii, jzi = ji * drri + jr * drii;
                const double wb = rri * d2, wd = an2 * rri * uri * d1;
                o0[6 * TSTR] = wb * jr; o1[6 * TSTR] = wb * ji;    // q2
                const double wc = wb * uri * ddri;
                o0[7 * TSTR] = wc * jr; o1[7 * TSTR] = wc * ji;    // q3
                o0[9 * TSTR] = wb * djr + wd * jzr; o1[9 * TSTR] = wb * dji + wd * jzi;  // q5
                if (!M0) {
                    const double wa = rri * (ssi * qmm) * d1;
                    o0[5 * TSTR] = wa * jr; o1[5 * TSTR] = wa * ji;    // q1
                    o0[8 * TSTR] = wa * djr; o1[8 * TSTR] = wa * dji;  // q4
                    const double dsi = si * qm * rri;
                    const double va = dsi * d1, vb = dsi * d2;
                    o0[10 * TSTR] = vb * jr; o1[10 * TSTR] = vb * ji;  // r1
                    o0[11 * TSTR] = va * jr; o1[11 * TSTR] = va * ji;  // r2
                    const double vf = an2 * va * uri;
                    o0[12 * TSTR] = vb * djr + vf * jzr; o1[12 * TSTR] = vb * dji + vf * jzi;  // r3
                    o0[13 * TSTR] = va * djr; o1[13 * TSTR] = va * dji;  // r4
                    const double ve = va * uri * ddri;
                    o0[14 * TSTR] = ve * djr; o1[14 * TSTR] = ve * dji;  // r5
                }
            }
        }
        ST_PROF(8)
        __syncthreads();
        ST_PROF(9)
#pragma unroll
        for (int ks = 0; ks < 2; ks++) {
            // fragment element of table t, rows [row0, row0 + 8): row0 + (lane >> 2), node 4 ks + (lane & 3)
            auto frag = [&](int t, int row0) -> double {
                const int r = row0 + fragrow;
                return op[(t * (2 * KR) + (r ^ (r >= KR ? 1 : 0))) * 8 + ((((ks ^ (r >> 1)) & 1) << 2) | fragk)];
            };
            double la[NTR][5];
#pragma unroll
            for (int r = 0; r < NTR; r++)
                if (tr0 + TRS * r < TMv) {
#pragma unroll
                    for (int t = 0; t < 5; t++) la[r][t] = frag(t, prow * KR + 8 * (tr0 + TRS * r));
                }
#pragma unroll
            for (int c = 0; c < TC; c++) {
                if (c >= TNv) continue;
                double rb[5];
                const int rbase = (M0 || even) ? 5 : 10;
#pragma unroll
                for (int t = 0; t < 5; t++) rb[t] = frag(rbase + t, pcol * KR + 8 * c);
                if (M0) {
                    // m = 0: dss = 0, the q1 and q4 terms vanish
#pragma unroll
                    for (int r = 0; r < NTR; r++)
                        if (tr0 + TRS * r < TMv) {
                            st_dmma(acc[r][c][0], la[r][1], rb[1]);
                            st_dmma(acc[r][c][1], la[r][3], rb[4]);
                            st_dmma(acc[r][c][0], la[r][4], rb[2]);
                        }
                } else if (even) {
#pragma unroll
                    for (int r = 0; r < NTR; r++)
                        if (tr0 + TRS * r < TMv) { st_dmma(acc[r][c][0], la[r][0], rb[0]); st_dmma(acc[r][c][1], la[r][2], rb[3]); }
#pragma unroll
                    for (int r = 0; r < NTR; r++)
                        if (tr0 + TRS * r < TMv) { st_dmma(acc[r][c][0], la[r][1], rb[1]); st_dmma(acc[r][c][1], la[r][3], rb[4]); }
#pragma unroll
                    for (int r = 0; r < NTR; r++)
                        if (tr0 + TRS * r < TMv) st_dmma(acc[r][c][0], la[r][4], rb[2]);
                } else {
#pragma unroll
                    for (int r = 0; r < NTR; r++)
                        if (tr0 + TRS * r < TMv) { st_dmma(acc[r][c][0], la[r][2], rb[0]); st_dmma(acc[r][c][1], la[r][0], rb[2]); }
#pragma unroll
                    for (int r = 0; r < NTR; r++)
                        if (tr0 + TRS * r < TMv) { st_dmma(acc[r][c][0], la[r][3], rb[1]); st_dmma(acc[r][c][1], la[r][1], rb[3]); }
#pragma unroll
                    for (int r = 0; r < NTR; r++)
                        if (tr0 + TRS * r < TMv) st_dmma(acc[r][c][1], la[r][4], rb[4]);
                }
            }
        }
        ST_PROF(10)
        __syncthreads();
        ST_PROF(11)
    }

    ST_PROF(1)
    // epilogue: lane holds (re, im) of X_F[a][b] with a = 4 tr + (lane >> 3), F = (lane >> 2) & 1, b = 4 tc + (lane & 3);
    // the y values come from lane ^ 4; lanes with F = 0 write the four system entries of the pair
    const int p = it.p;
    const double PI = 3.14159265358979323846;
    const double kw = PI * 2.0 / a.lam[p];
    const double ppi = kw * kw, pir = ppi * a.mr[p], pii = ppi * a.mi[p];
    // NM <= 32: the block is the whole mode; the systems are combined in the (now free) operand area, solved here and only
    // the T block leaves the SM.  Larger modes write their part of [Q | RgQ] to global memory for k_st_solve.
    const bool fused = NM <= KR && (a.fuse & 1);
    double2 *sys = fused ? reinterpret_cast<double2 *>(op) : a.sys + it.sys_off + st_sys_mode_off(nmax, m);
#pragma unroll
    for (int r = 0; r < NTR; r++)
#pragma unroll
        for (int c = 0; c < TC; c++) {
            const bool tv = (tr0 + TRS * r < TMv) && (c < TNv);  // warp-uniform
            if (!tv) continue;
            double xa[4], xb[4];
            xa[0] = acc[r][c][0][0]; xa[1] = acc[r][c][0][1];
            xb[0] = acc[r][c][1][0]; xb[1] = acc[r][c][1][1];
            xa[2] = __shfl_xor_sync(0xffffffffu, xa[0], 4); xa[3] = __shfl_xor_sync(0xffffffffu, xa[1], 4);
            xb[2] = __shfl_xor_sync(0xffffffffu, xb[0], 4); xb[3] = __shfl_xor_sync(0xffffffffu, xb[1], 4);
            const int al = 4 * (tr0 + TRS * r) + (lane >> 3), bl = 4 * c + (lane & 3);
            if (((lane >> 2) & 1) == 0 && al < hrow && bl < hcol) {
                const int kk1 = kr0 + 2 * al + prow, kk2 = kc0 + 2 * bl + pcol;
                if (even) store_even(sys, NM, nmin, kk1, kk2, ppi, pir, pii, a.dd, xa, xa + 2, xb, xb + 2);
                else store_odd(sys, NM, nmin, kk1, kk2, ppi, pir, pii, a.dd, xa, xa + 2, xb, xb + 2);
            }
        }
    if (M0) {
        // opposite-parity pairs vanish identically for m = 0
        const int ldc = 2 * NM;
        const double2 z = make_double2(0.0, 0.0);
        for (int t = tid; t < KR * KR; t += THREADS) {
            const int q1 = kr0 + t / KR, q2 = kc0 + t % KR;
            if (q1 >= NM || q2 >= NM || (((q1 + q2) & 1) == 0)) continue;
            for (int cl = 0; cl < 2; cl++) {
                sys[(long)(cl * NM + q2) * ldc + q1] = z;
                sys[(long)(cl * NM + q2) * ldc + NM + q1] = z;
            }
        }
    }
    if (!fused) return;
#if ST_LEGACY_FUSED_SOLVERS
    __shared__ unsigned long long keyslot[4][2];
    __shared__ double2 s_colbuf[4][2 * 32], s_rowbuf[THREADS / 32][16];  // register solver: published pivot columns / rows
#endif
    // register-tiled solver (hs_tilesolve.cuh): m >= 1: two systems of KR rows on half the CTA each; m = 0: four half-size systems
    typedef TileSolve<KR, KR / 8, 4, KR / 8> TS1;
    typedef TileSolve<KR / 2, KR / 16, 4, KR / 16> TS0;
    __shared__ typename TS1::Buf s_tb1[M0 ? 1 : 2];  // (the compiler drops the array the instantiation does not use)
    __shared__ typename TS0::Buf s_tb0[M0 ? 4 : 1];
    __shared__ int sing;
    if (tid == 0) sing = 0;
    __syncthreads();
    ST_PROF(2)
    const int ldc = 2 * NM;
    constexpr int LM0 = THREADS / 4, LM1 = THREADS / 2;  // lanes per system: m = 0 four half-size systems, m >= 1 two systems
    if (M0) {
        // four half-size systems (T12 = T21 = 0)
        const int q = tid / LM0;
        SysDesc d;
        d.base = sys + (long)(q >> 1) * NM * ldc; d.ld = ldc; d.off = q & 1; d.stride = 2; d.rhs = NM; d.n = (NM + 1 - (q & 1)) >> 1;
        if (d.n > 0) {
#if ST_LEGACY_FUSED_SOLVERS
            if (a.fuse & ST_FUSE_TILE) TS0::solve(d, tid % LM0, 1 + q, &s_tb0[M0 ? q : 0], &sing);
            else if ((a.fuse & 8) && KR == 32) solve_rows16(d, (tid % LM0) >> 5, LM0 / 32, lane, 1 + q, s_colbuf[q], s_rowbuf[tid >> 5], &sing);
            else if (a.fuse & 2) solve_grp2s(d, tid % LM0, LM0, 1 + q, keyslot[q], &sing);
            else solve_grp(d, tid % LM0, LM0, 1 + q, keyslot[q], &sing);
#else
            TS0::solve(d, tid % LM0, 1 + q, &s_tb0[M0 ? q : 0], &sing);
#endif
        }
    } else {
        // the two parity classes
        const int q = tid / LM1;
        SysDesc d;
        d.base = sys + (long)q * NM * ldc; d.n = NM; d.ld = ldc; d.off = 0; d.stride = 1; d.rhs = NM;
        // measured per elimination step inside the pipeline (W2): wide modes (NM 17..32, four warps per system) 2 940 -> 2 260 cycles;
        // no gain for the narrow variant and the m = 0 half systems, which therefore keep the shared-memory elimination
#if ST_LEGACY_FUSED_SOLVERS
        if (a.fuse & ST_FUSE_TILE) TS1::solve(d, tid % LM1, 1 + q, &s_tb1[M0 ? 0 : q], &sing);
        else if ((a.fuse & 4) && (KR == 32 || (a.fuse & 8))) solve_rows16(d, (tid % LM1) >> 5, LM1 / 32, lane, 1 + q, s_colbuf[q], s_rowbuf[tid >> 5], &sing);
        else if (a.fuse & 2) solve_grp2s(d, tid % LM1, LM1, 1 + q, keyslot[q], &sing);
        else solve_grp(d, tid % LM1, LM1, 1 + q, keyslot[q], &sing);
#else
        TS1::solve(d, tid % LM1, 1 + q, &s_tb1[M0 ? 0 : q], &sing);
#endif
    }
    __syncthreads();
    ST_PROF(3)
    const long long toff = it.t_off;
    if (toff >= 0) {
        double2 *tg = reinterpret_cast<double2 *>(a.tarena) + toff + st_t_mode_off(nmax, m);
        for (int t = tid; t < 2 * NM * NM; t += THREADS) {
            const int row = t / NM, col = t - row * NM;
            tg[t] = sys[(long)row * ldc + NM + col];
        }
    }
    if (tid == 0) {
        if (M0) {
            double qext = 0.0, qsca = 0.0;
            for (int n_ = 1; n_ <= nmax; n_++) {
                const int k = n_ - 1;
                const double2 t1 = sys[(long)(((n_ + 1) & 1) * NM + k) * ldc + NM + k];
                const double2 t2 = sys[(long)((n_ & 1) * NM + k) * ldc + NM + k];
                const double dn1 = (double)(2 * n_ + 1);
                qsca += dn1 * (t1.x * t1.x + t1.y * t1.y + t2.x * t2.x + t2.y * t2.y);
                qext += (t1.x + t2.x) * dn1;
            }
            a.res[mi.item].qsca = qsca;
            a.res[mi.item].qext = qext;
        }
        if (sing) a.res[mi.item].sing = 1;
    }
    ST_PROF(4)
    if (a.prof && threadIdx.x == 0) { atomicAdd(&a.prof[5], 1ull); atomicAdd(&a.prof[6], (unsigned long long)NM); atomicAdd(&a.prof[7], (unsigned long long)g); atomicAdd(&a.prof[12], (unsigned long long)(M0 ? (NM + 1) / 2 : NM)); }
#undef ST_PROF
}

template <bool M0, int KR>
__global__ void __launch_bounds__(8 * KR, KR == 32 ? (M0 ? ST_WIDE0_CTAS : 2) : 4) k_st_assemble(StArgs a, const StATask *tasks, int ntasks)
{
    extern __shared__ __align__(16) double op[];
    if ((int)blockIdx.x >= ntasks) return;
    st_assemble_task<M0, KR>(a, tasks[blockIdx.x], op);
}

// resident CTAs fetching tasks from the device-built list (see k_st_radial_q)
template <bool M0, int KR>
__global__ void __launch_bounds__(8 * KR, KR == 32 ? (M0 ? ST_WIDE0_CTAS : 2) : 4) k_st_assemble_q(StArgs a, const StATask *tasks, StQueue *q)
{
    extern __shared__ __align__(16) double op[];
    __shared__ int s_next;
    // thread 0 fetches the index of the task after this one while the CTA works (the cursor round trip is off the path)
    int nxt = 0;
    if (threadIdx.x == 0) { nxt = atomicAdd(&q->cursor, 1); if (nxt >= q->count) nxt = -1; }
    for (;;) {
        if (threadIdx.x == 0) s_next = nxt;
        __syncthreads();
        const int t = s_next;
        if (t < 0) return;
        if (threadIdx.x == 0) { nxt = atomicAdd(&q->cursor, 1); if (nxt >= q->count) nxt = -1; }
        st_assemble_task<M0, KR>(a, tasks[t], op);
        __syncthreads();  // the operand area / fused systems and s_next are re-used by the next task
    }
}

// ---- G: parity-class systems of one (item, mode): T = -RgQ Q^-1 by the lock-step Gauss-Jordan of hs_tmat.cuh ----
template <int THREADS, bool GLOBAL>
__device__ __forceinline__ void st_solve_task(const StArgs &a, const int mix, double2 *ssys)
{
    __shared__ SysDesc sd[HS_MAXSYS];
    __shared__ SolveCtl sc;
    __shared__ int sing;
    const StModeItem mi = st_unpack_mode(mix);
    const StItem it = a.items[mi.item];
    const int m = mi.m, nmax = it.nmax;
    const int nmin = m > 1 ? m : 1, NM = nmax - nmin + 1, ldc = 2 * NM;
    double2 *gsys = a.sys + it.sys_off + st_sys_mode_off(nmax, m);
    const int tot = 4 * NM * NM;
    double2 *wk = GLOBAL ? gsys : ssys;
    if (!GLOBAL)
        for (int t = threadIdx.x; t < tot; t += THREADS) wk[t] = gsys[t];
    int nsys;
    if (m == 0) {
        // T12 = T21 = 0: each parity class splits into two interleaved half-size systems
        nsys = 0;
        for (int cl = 0; cl < 2; cl++)
            for (int par = 0; par < 2; par++) {
                const int nn = (NM + 1 - par) >> 1;
                if (nn > 0) {
                    if (threadIdx.x == 0) {
                        SysDesc d;
                        d.base = wk + (long)cl * NM * ldc; d.ld = ldc; d.off = par; d.stride = 2; d.rhs = NM; d.n = nn;
                        sd[nsys] = d;
                    }
                    nsys++;
                }
            }
    } else {
        nsys = 2;
        if (threadIdx.x < 2) {
            SysDesc d;
            d.base = wk + (long)threadIdx.x * NM * ldc; d.n = NM; d.ld = ldc; d.off = 0; d.stride = 1; d.rhs = NM;
            sd[threadIdx.x] = d;
        }
    }
    if (threadIdx.x == 0) sing = 0;
    __syncthreads();
    if (GLOBAL) {
        Grp<false> grp;
        grp.tid = threadIdx.x; grp.nt = THREADS;
        solve_multi(grp, sd, nsys, &sc, &sing);
    } else {
        if (THREADS == 256 && NM <= 64) {
            // as in the fused path: m = 0: four half-size systems on 64 lanes each, m >= 1: two systems on 128 lanes each
            // (a lane pair owns up to two columns: 2 NM <= 128 resp. NM + 1 <= 64)
            __shared__ unsigned long long keyslot[4][2];
            const int warp = threadIdx.x >> 5;
            const int q = (m == 0) ? (warp >> 1) : (warp >> 2);
            if (q < nsys && sd[q].n > 0) solve_grp2s(sd[q], threadIdx.x & (m == 0 ? 63 : 127), m == 0 ? 64 : 128, 1 + q, keyslot[q], &sing);
        } else if (THREADS == 64) {
            // NM <= 16: one warp per system (a lane pair owns up to two of the 2 n <= 32 columns); m = 0: two systems per warp, one after the other
            __shared__ unsigned long long keyslot[4][2];
            const int warp = threadIdx.x >> 5;
            for (int q = warp; q < nsys; q += 2)
                if (sd[q].n > 0) solve_grp2s(sd[q], threadIdx.x & 31, 32, 1 + warp, keyslot[q], &sing);
        } else {
            // one warp per system (m = 0: four half-size systems, m >= 1: the two parity classes)
            const int w = threadIdx.x >> 5;
            if (w < nsys) solve_sub(sd[w], threadIdx.x & 31, 32, sd[w].n, &sing);
        }
        __syncthreads();
    }
    const long long toff = it.t_off;
    if (toff >= 0) {
        double2 *tg = reinterpret_cast<double2 *>(a.tarena) + toff + st_t_mode_off(nmax, m);
        for (int t = threadIdx.x; t < 2 * NM * NM; t += THREADS) {
            const int row = t / NM, col = t - row * NM;
            tg[t] = wk[(long)row * ldc + NM + col];
        }
    }
    if (threadIdx.x == 0) {
        if (m == 0) {
            double qext = 0.0, qsca = 0.0;
            for (int n_ = 1; n_ <= nmax; n_++) {
                const int k = n_ - 1;
                const double2 t1 = wk[(long)(((n_ + 1) & 1) * NM + k) * ldc + NM + k];
                const double2 t2 = wk[(long)((n_ & 1) * NM + k) * ldc + NM + k];
                const double dn1 = (double)(2 * n_ + 1);
                qsca += dn1 * (t1.x * t1.x + t1.y * t1.y + t2.x * t2.x + t2.y * t2.y);
                qext += (t1.x + t2.x) * dn1;
            }
            a.res[mi.item].qsca = qsca;
            a.res[mi.item].qext = qext;
        }
        if (sing) a.res[mi.item].sing = 1;
    }
}

template <int THREADS, bool GLOBAL>
__global__ void __launch_bounds__(THREADS) k_st_solve(StArgs a, const int *mlist, int n)
{
    extern __shared__ __align__(16) double2 ssys[];
    if ((int)blockIdx.x >= n) return;
    st_solve_task<THREADS, GLOBAL>(a, mlist[blockIdx.x], ssys);
}

template <int THREADS, bool GLOBAL>
__global__ void __launch_bounds__(THREADS) k_st_solve_q(StArgs a, const int *mlist, StQueue *q)
{
    extern __shared__ __align__(16) double2 ssys[];
    __shared__ int s_next;
    const int n = q->count;
    for (;;) {
        if (threadIdx.x == 0) s_next = atomicAdd(&q->cursor, 1);
        __syncthreads();
        const int t = s_next;
        if (t >= n) return;
        st_solve_task<THREADS, GLOBAL>(a, mlist[t], ssys);
        __syncthreads();
    }
}


// ---- A + G for small blocks (NM <= 8): one warp per (item, mode) ----
// The four parity blocks are single 8 x 8 DMMA tiles, so a CTA-wide kernel would be all padding and barriers.  Here a
// warp builds its own operand rows for 4 nodes at a time ([15 tables][16 rows][4 nodes], private shared-memory slice),
// contracts them, combines the accumulators into the two parity-class systems in the same slice, solves them with the
// warp-wide lock-step Gauss-Jordan and writes the T block / (Qsca, Qext): no CTA barrier, no global round trip of Q.
#define ST_SMALL_NM 8
#ifndef ST_SMALL_CTAS
#define ST_SMALL_CTAS 2
#endif
#define ST_SMALL_SLICE (15 * 16 * 4)  // doubles per warp

// one (item, mode) with NM <= 8 on one warp; op = the warp's private slice, sd / sing = its descriptor slots
__device__ __forceinline__ void st_small_task(const StArgs &a, const int mix, double *op, SysDesc *sd, int *sing, int lane)
{
    long long tprof = a.prof ? clock64() : 0;
#define SM_PROF(slot) if (a.prof && lane == 0) { const long long now_ = clock64(); atomicAdd(&a.prof[slot], (unsigned long long)(now_ - tprof)); tprof = now_; }
    const StModeItem mi = st_unpack_mode(mix);
    const StItem it = a.items[mi.item];
    const int m = mi.m, nmax = it.nmax, g = it.g, gp = it.gp;
    const int nmin = m > 1 ? m : 1, NM = nmax - nmin + 1;
    const int h0 = (NM + 1) >> 1, h1 = NM >> 1;
    const double *nod = a.tab + it.tab_off;
    const double *tb = nod + (long)NA_COUNT * gp;
    const long tsz = (long)nmax * gp;
    const double *w1 = a.wig + it.wig_off + st_wig_mode_off(nmax, m, gp);
    const double *w2 = w1 + (long)NM * gp;
    const double qm = (double)m, qmm = qm * qm;
    const int kk = lane >> 2, ii = lane & 3;
    const int rho = (kk & 1) * 8 + 2 * (kk >> 1);
    // physical row = r ^ (r >= 8): a half-warp stores the row pairs of two even and two odd k; with the odd block's pairs
    // swapped its 16 lanes cover the 16 bank pairs once (see k_st_assemble)
    const int rsw = kk & 1;
    double *o0 = op + (rho ^ rsw) * 4 + ii, *o1 = op + ((rho + 1) ^ rsw) * 4 + ii;
    const int n1 = nmin + kk;
    const double an1 = (double)(n1 * (n1 + 1));
    const int fragrow = lane >> 2, fragk = lane & 3;
    double acc[4][2][2];  // [EE, OO, EO, OE][A, B][re, im]
#pragma unroll
    for (int b = 0; b < 4; b++)
#pragma unroll
        for (int e = 0; e < 4; e++) acc[b][e >> 1][e & 1] = 0.0;
    // the table values of the next 4 nodes are loaded (into registers) before the current 4 are processed: the L2 latency
    // of a chunk is covered by the build + contraction of the previous one
    struct Vals { double d1, d2, fj, fdj, fy, fdy, jr, ji, djr, dji, rri, uri, ddri, drri, drii, ssi, si; };
    auto load_vals = [&](int i0) -> Vals {
        Vals v = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        const int i = i0 + ii;
        if (i < g && kk < NM) {
            const long o = (long)(n1 - 1) * gp + i;
            v.d1 = w1[(long)kk * gp + i]; v.d2 = w2[(long)kk * gp + i];
            v.fj = tb[o]; v.fdj = tb[tsz + o]; v.fy = tb[2 * tsz + o]; v.fdy = tb[3 * tsz + o];
            v.jr = tb[4 * tsz + o]; v.ji = tb[5 * tsz + o]; v.djr = tb[6 * tsz + o]; v.dji = tb[7 * tsz + o];
            v.rri = nod[NA_RR * gp + i]; v.uri = nod[NA_DR * gp + i]; v.ddri = nod[NA_DDR * gp + i];
            v.drri = nod[NA_DRR * gp + i]; v.drii = nod[NA_DRI * gp + i];
            if (m) { v.ssi = nod[NA_SS * gp + i]; v.si = nod[NA_S * gp + i]; }
        }
        return v;
    };
    Vals nx = load_vals(0);
    SM_PROF(16)
    for (int i0 = 0; i0 < g; i0 += 4) {
        {
            const Vals cv = nx;
            if (i0 + 4 < g) nx = load_vals(i0 + 4);
            const double d1 = cv.d1, d2 = cv.d2, fj = cv.fj, fdj = cv.fdj, fy = cv.fy, fdy = cv.fdy, jr = cv.jr, ji = cv.ji, djr = cv.djr, dji = cv.dji;
            const double rri = cv.rri, uri = cv.uri, ddri = cv.ddri, drri = cv.drri, drii = cv.drii, ssi = cv.ssi, si = cv.si;
            __syncwarp();  // previous contraction has read the slice
            o0[0 * 64] = d1 * fdj; o1[0 * 64] = d1 * fdy;
            o0[1 * 64] = d2 * fdj; o1[1 * 64] = d2 * fdy;
            const double p3j = d1 * fj, p3y = d1 * fy;
            o0[2 * 64] = p3j; o1[2 * 64] = p3y;
            o0[3 * 64] = d2 * fj; o1[3 * 64] = d2 * fy;
            o0[4 * 64] = an1 * p3j; o1[4 * 64] = an1 * p3y;
            const double jzr = jr * drri - ji * drii, jzi = ji * drri + jr * drii;
            const double wb = rri * d2, wd = an1 * rri * uri * d1;
            const double wa = rri * (ssi * qmm) * d1;
            o0[5 * 64] = wa * jr; o1[5 * 64] = wa * ji;
            o0[6 * 64] = wb * jr; o1[6 * 64] = wb * ji;
            const double wc = wb * uri * ddri;
            o0[7 * 64] = wc * jr; o1[7 * 64] = wc * ji;
            o0[8 * 64] = wa * djr; o1[8 * 64] = wa * dji;
            o0[9 * 64] = wb * djr + wd * jzr; o1[9 * 64] = wb * dji + wd * jzi;
            if (m) {
                const double dsi = si * qm * rri;
                const double va = dsi * d1, vb = dsi * d2;
                o0[10 * 64] = vb * jr; o1[10 * 64] = vb * ji;
                o0[11 * 64] = va * jr; o1[11 * 64] = va * ji;
                const double vf = an1 * va * uri;
                o0[12 * 64] = vb * djr + vf * jzr; o1[12 * 64] = vb * dji + vf * jzi;
                o0[13 * 64] = va * djr; o1[13 * 64] = va * dji;
                const double ve = va * uri * ddri;
                o0[14 * 64] = ve * djr; o1[14 * 64] = ve * dji;
            }
            __syncwarp();
        }
        auto frag = [&](int t, int row0) -> double { return op[(t * 16 + ((row0 + fragrow) ^ (row0 >> 3))) * 4 + fragk]; };
        double lE[5], lO[5];
#pragma unroll
        for (int t = 0; t < 5; t++) { lE[t] = frag(t, 0); lO[t] = frag(t, 8); }
        {
            double rE[5], rO[5];
#pragma unroll
            for (int t = 0; t < 5; t++) { rE[t] = frag(5 + t, 0); rO[t] = frag(5 + t, 8); }
            if (m) {
                st_dmma(acc[0][0], lE[0], rE[0]); st_dmma(acc[1][0], lO[0], rO[0]);
                st_dmma(acc[0][1], lE[2], rE[3]); st_dmma(acc[1][1], lO[2], rO[3]);
            }
            st_dmma(acc[0][0], lE[1], rE[1]); st_dmma(acc[1][0], lO[1], rO[1]);
            st_dmma(acc[0][1], lE[3], rE[4]); st_dmma(acc[1][1], lO[3], rO[4]);
            st_dmma(acc[0][0], lE[4], rE[2]); st_dmma(acc[1][0], lO[4], rO[2]);
        }
        if (m) {
            double rE[5], rO[5];
#pragma unroll
            for (int t = 0; t < 5; t++) { rE[t] = frag(10 + t, 0); rO[t] = frag(10 + t, 8); }
            st_dmma(acc[2][0], lE[2], rO[0]); st_dmma(acc[3][0], lO[2], rE[0]);
            st_dmma(acc[2][1], lE[0], rO[2]); st_dmma(acc[3][1], lO[0], rE[2]);
            st_dmma(acc[2][0], lE[3], rO[1]); st_dmma(acc[3][0], lO[3], rE[1]);
            st_dmma(acc[2][1], lE[1], rO[3]); st_dmma(acc[3][1], lO[1], rE[3]);
            st_dmma(acc[2][1], lE[4], rO[4]); st_dmma(acc[3][1], lO[4], rE[4]);
        }
    }
    __syncwarp();
    SM_PROF(17)
    // combine into the parity-class systems, stored in the same slice ([2][NM][2 NM] complex <= 512 doubles)
    const int p = it.p;
    const double PI = 3.14159265358979323846;
    const double kw = PI * 2.0 / a.lam[p];
    const double ppi = kw * kw, pir = ppi * a.mr[p], pii = ppi * a.mi[p];
    double2 *sys = reinterpret_cast<double2 *>(op);
    const int ldc = 2 * NM;
    if (m == 0) {
        const double2 z = make_double2(0.0, 0.0);
        for (int t = lane; t < NM * NM; t += 32) {
            const int q1 = t / NM, q2 = t - q1 * NM;
            if (((q1 + q2) & 1) == 0) continue;
            for (int cl = 0; cl < 2; cl++) { sys[(cl * NM + q2) * ldc + q1] = z; sys[(cl * NM + q2) * ldc + NM + q1] = z; }
        }
    }
#pragma unroll
    for (int b = 0; b < 4; b++) {
        if (b >= 2 && m == 0) break;
        const int prow = (b == 1 || b == 3) ? 1 : 0, pcol = (b == 1 || b == 2) ? 1 : 0;
        const int hrow = prow ? h1 : h0, hcol = pcol ? h1 : h0;
        double xa[4], xb[4];
        xa[0] = acc[b][0][0]; xa[1] = acc[b][0][1]; xb[0] = acc[b][1][0]; xb[1] = acc[b][1][1];
        xa[2] = __shfl_xor_sync(0xffffffffu, xa[0], 4); xa[3] = __shfl_xor_sync(0xffffffffu, xa[1], 4);
        xb[2] = __shfl_xor_sync(0xffffffffu, xb[0], 4); xb[3] = __shfl_xor_sync(0xffffffffu, xb[1], 4);
        const int al = lane >> 3, bl = lane & 3;
        if (((lane >> 2) & 1) == 0 && al < hrow && bl < hcol) {
            const int kk1 = 2 * al + prow, kk2 = 2 * bl + pcol;
            if (b < 2) store_even(sys, NM, nmin, kk1, kk2, ppi, pir, pii, a.dd, xa, xa + 2, xb, xb + 2);
            else store_odd(sys, NM, nmin, kk1, kk2, ppi, pir, pii, a.dd, xa, xa + 2, xb, xb + 2);
        }
    }
    // solve
    SM_PROF(18)
    int nsys;
    if (m == 0) {
        nsys = 0;
        for (int cl = 0; cl < 2; cl++)
            for (int par = 0; par < 2; par++) {
                const int nn = (NM + 1 - par) >> 1;
                if (nn > 0) {
                    if (lane == 0) {
                        SysDesc d;
                        d.base = sys + cl * NM * ldc; d.ld = ldc; d.off = par; d.stride = 2; d.rhs = NM; d.n = nn;
                        sd[nsys] = d;
                    }
                    nsys++;
                }
            }
    } else {
        nsys = 2;
        if (lane < 2) {
            SysDesc d;
            d.base = sys + lane * NM * ldc; d.n = NM; d.ld = ldc; d.off = 0; d.stride = 1; d.rhs = NM;
            sd[lane] = d;
        }
    }
    if (lane == 0) *sing = 0;
    __syncwarp();
    {
        // m = 0: four half-size systems on 8 lanes each; m >= 1: the two parity classes on 16 lanes each
        const int gw = m == 0 ? 8 : 16, grp_ = lane / gw, gl = lane - grp_ * gw;
        SysDesc d;
        if (grp_ < nsys) d = sd[grp_];
        else { d = sd[0]; d.n = 0; }
        solve_subs(d, gl, gw, m == 0 ? ((NM + 1) >> 1) : NM, sing);
    }
    SM_PROF(19)
    const long long toff = it.t_off;
    if (toff >= 0) {
        double2 *tg = reinterpret_cast<double2 *>(a.tarena) + toff + st_t_mode_off(nmax, m);
        for (int t = lane; t < 2 * NM * NM; t += 32) {
            const int row = t / NM, col = t - row * NM;
            tg[t] = sys[row * ldc + NM + col];
        }
    }
    if (lane == 0) {
        if (m == 0) {
            double qext = 0.0, qsca = 0.0;
            for (int n_ = 1; n_ <= nmax; n_++) {
                const int k = n_ - 1;
                const double2 t1 = sys[(((n_ + 1) & 1) * NM + k) * ldc + NM + k];
                const double2 t2 = sys[((n_ & 1) * NM + k) * ldc + NM + k];
                const double dn1 = (double)(2 * n_ + 1);
                qsca += dn1 * (t1.x * t1.x + t1.y * t1.y + t2.x * t2.x + t2.y * t2.y);
                qext += (t1.x + t2.x) * dn1;
            }
            a.res[mi.item].qsca = qsca;
            a.res[mi.item].qext = qext;
        }
        if (*sing) a.res[mi.item].sing = 1;
    }
    SM_PROF(20)
    if (a.prof && lane == 0) { atomicAdd(&a.prof[21], 1ull); atomicAdd(&a.prof[22], (unsigned long long)NM); atomicAdd(&a.prof[23], (unsigned long long)g); }
#undef SM_PROF
}

__global__ void __launch_bounds__(256, ST_SMALL_CTAS) k_st_mode_small(StArgs a, const int *mlist, int n)
{
    extern __shared__ __align__(16) double smraw[];
    __shared__ SysDesc sds[8][4];
    __shared__ int sings[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wi = blockIdx.x * 8 + warp;
    if (wi >= n) return;
    st_small_task(a, mlist[wi], smraw + warp * ST_SMALL_SLICE, sds[warp], &sings[warp], lane);
}

// resident warps fetching (item, mode) pairs from the device-built list (see k_st_radial_q)
__global__ void __launch_bounds__(256, ST_SMALL_CTAS) k_st_mode_small_q(StArgs a, const int *mlist, StQueue *q)
{
    extern __shared__ __align__(16) double smraw[];
    __shared__ SysDesc sds[8][4];
    __shared__ int sings[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = q->count;
    // four (item, mode) pairs per cursor round trip when the list is long; a short list (the latency-bound tail rounds, or a
    // small batch of particles with many quadrature nodes) is dealt one pair at a time so that no warp queues up work
    const int batch = n >= 16 * (int)(gridDim.x * 8) ? 4 : 1;
    for (;;) {
        int t0 = 0;
        if (lane == 0) t0 = atomicAdd(&q->cursor, batch);
        t0 = __shfl_sync(0xffffffffu, t0, 0);
        if (t0 >= n) return;
        const int t1 = min(t0 + batch, n);
        for (int t = t0; t < t1; t++) {
            st_small_task(a, mlist[t], smraw + warp * ST_SMALL_SLICE, sds[warp], &sings[warp], lane);
            __syncwarp();
        }
    }
}

#endif  // HS_HOST_EMU

}  // namespace hs
