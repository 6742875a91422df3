// hs_dvctl.h -- the NMAX / NGAUSS convergence controller of `calctmat` and the planner of a
// particle's next round, as plain functions shared by the device-driven rounds (hs_staged_dev.h: k_dv_plan, k_dv_fill) and the
// single-lane host emulation (tests/emu): no CUDA intrinsics in here.
//
// Round structure: a particle in phase 0 (NMAX search) runs kspec speculative trials NMAX, NMAX + 1, ... (NGAUSS = NDGS NMAX) per
// round; the controller consumes their (Qsca, Qext) in the reference's order and stops at the first one that passes the
// reference's test (|dQsca / Qsca| <= DDELT and |dQext / Qext| <= DDELT against the previous trial) -- trials speculated beyond it
// are discarded unread.  Phase 1 (NGAUSS search): one trial NGAUSS + 1 per round, launched together with every mode m >= 1 of that
// configuration; phase 2: NGAUSS hit NPNG1 while NMAX converged, the modes of that configuration only.
#pragma once
#include "hs_compat.cuh"
#include "../../include/hydrotm.h"

namespace hs {

struct DvP {  // controller state of one particle (StP of the host-driven path)
    int idx, phase, nmax, g, ntr, status, done, sing;
    double q1s, q1e;
    long long t_off;
    int first_item, n_items, kspec, written;
    double dprev, dlast;
};

struct DvTrial { double qsca, qext; long long t_off; int sing, g; };  // what a finished item reports (+ its NGAUSS)

// start of a call: starting order s0 (< 0: bad input, -(n0 + 1)), lower-bound guess lb of the converged order
__host__ __device__ inline void dv_init(DvP &s, int idx, int s0, int lb, int spec0)
{
    s.idx = idx; s.phase = 0; s.nmax = s0; s.g = 0; s.ntr = 0; s.status = HS_ST_OK; s.done = 0; s.sing = 0;
    s.q1s = 0.0; s.q1e = 0.0; s.t_off = -1; s.first_item = 0; s.n_items = 0; s.written = 0;
    // every trial from the starting order to the converged one is needed, so a lower bound of the converged order is free
    // parallelism: the first round runs them all at once
    const int a = lb - (s0 > 0 ? s0 : 0) + 1 + spec0, b = 12 + spec0;
    s.kspec = a < b ? a : b;
    if (s.kspec < 2) s.kspec = 2;
    s.dprev = 0.0; s.dlast = 0.0;
    if (s.nmax < 0) { s.done = 2; s.status = HS_ST_BAD_INPUT; s.nmax = -s.nmax - 1; }
    else if (s.nmax >= HS_NPN1) { s.done = 2; s.status = HS_ST_NMAX_LIMIT; }
}

// items of particle `s` in the coming round: calls f(nmax, g, nmodes) for each; returns false (and *gmiss = the rule) if a
// Gauss rule the round needs is not in the tables yet (the particle then waits for the host to extend them)
template <class F>
__host__ __device__ inline bool dv_items_of(const DvP &s, int ndgs, int spec, int gmax, int gfmax, int np, int *gmiss, F f)
{
    const int cnt = s.phase == 0 ? (spec > 0 ? spec : s.kspec) : 1;
    for (int q = 0; q < cnt; q++) {
        const int nm = s.phase == 0 ? s.nmax + q : s.nmax;
        const int g = s.phase == 0 ? nm * ndgs : (s.phase == 1 ? s.g + 1 : s.g);
        if (nm > HS_NPN1 || g > HS_NPNG1) break;
        if (g > gmax || (np == -2 && g / 2 + 1 > gfmax)) { if (gmiss) *gmiss = g; return false; }
    }
    for (int q = 0; q < cnt; q++) {
        const int nm = s.phase == 0 ? s.nmax + q : s.nmax;
        const int g = s.phase == 0 ? nm * ndgs : (s.phase == 1 ? s.g + 1 : s.g);
        if (nm > HS_NPN1 || g > HS_NPNG1) break;  // the controller reports the limit when it gets there
        f(nm, g, s.phase == 0 ? 1 : nm + 1);
    }
    return true;
}

// controller over the n_items results of the round just finished (get(q) = q-th item of the particle, in plan order), then
// the limit test of the coming round.  A port of StGroup::finish (hs_staged_host.h): same comparisons, same order.
template <class GET>
__host__ __device__ inline void dv_consume(DvP &s, double ddelt, int ndgs, int margin, GET get)
{
    if (!s.done && s.n_items > 0) {
        for (int q = 0; q < s.n_items; q++) {
            const DvTrial r = get(q);
            if (s.phase == 2) {  // modes of a configuration whose convergence is already decided
                if (r.t_off == -2) { s.status = HS_ST_ARENA_FULL; s.done = 2; break; }
                s.t_off = r.t_off; s.sing |= r.sing; s.done = 1;
                break;
            }
            const double dsca = fabs((s.q1s - r.qsca) / r.qsca), dext = fabs((s.q1e - r.qext) / r.qext);
            s.q1s = r.qsca; s.q1e = r.qext;
            s.ntr++;
            s.g = r.g;
            s.sing |= r.sing;
            const bool ok = (dsca <= ddelt && dext <= ddelt);
            if (s.phase == 0) {
                s.dprev = s.dlast; s.dlast = fmax(dsca, dext) / ddelt;
                if (ok) { s.phase = (r.g == HS_NPNG1) ? 2 : 1; break; }
                else if (s.nmax + 1 > HS_NPN1) { s.status = HS_ST_NMAX_LIMIT; s.done = 2; break; }
                else s.nmax = s.nmax + 1;
            } else {
                if (r.t_off == -2) { s.status = HS_ST_ARENA_FULL; s.done = 2; break; }
                s.t_off = r.t_off;
                if (ok) s.done = 1;
                else if (r.g == HS_NPNG1) { s.status = HS_ST_NGAUSS_LIMIT; s.done = 1; }
                break;
            }
        }
        if (!s.done && s.phase == 0) {
            // speculation depth of the next round from the convergence trend: the relative change of Qsca / Qext shrinks
            // roughly geometrically with NMAX, so log(d_last) / log(d_prev / d_last) more trials bring it under DDELT
            int k = (s.dlast > 0.0 && s.dlast < 100.0) ? 1 : 2;
            if (s.dprev > s.dlast && s.dlast > 1.0 && isfinite(s.dprev)) {
                const double r = log(s.dlast) / log(s.dprev / s.dlast);
                k = (int)fmin(8.0, fmax(1.0, ceil(r) + (double)margin));
            }
            s.kspec = k;
        }
    }
    s.n_items = 0;
    if (!s.done && s.phase == 0 && (long long)s.nmax * ndgs > HS_NPNG1) { s.status = HS_ST_NGAUSS_LIMIT; s.done = 2; }
}

}  // namespace hs
