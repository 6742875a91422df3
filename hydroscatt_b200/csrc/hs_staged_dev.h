// hs_staged_dev.h -- device-driven rounds of the staged T-matrix pipeline (hs_staged.cuh).
//
// Same stage kernels and the same NMAX / NGAUSS convergence controller of `calctmat` as
// hs_staged_host.h, but the controller, the planner of the next round and the task lists live on the device: a call is a
// fixed sequence of launches with no host round trip between rounds.
//
//   round r of a group = k_dv_fill   thread = particle: items, (item, mode) list, radial / assembly / solve task lists,
//                                    arena slots of final items, cleared result records
//                        S, W        k_st_radial_q, k_st_wigner_q: resident warps, static stride over the device-built lists
//                        A, G        k_st_assemble_q, k_st_mode_small_q, k_st_solve_q: resident CTAs / warps fetch tasks from a
//                                    cursor (the next index is fetched while the current task runs)
//                        k_dv_plan   thread = particle: controller over the results of round r (hs_dvctl.h: the reference's
//                                    order), finished particles write their output record, the others count the work of
//                                    round r + 1; block-wide prefix of the counts
//                        k_dv_scan   one CTA: prefix over the block totals = list offsets, capacity cut, queue lengths;
//                                    the header also goes to pinned host memory
//
// The host enqueues the rounds of every group up front (the number the previous call needed), waits once, looks at the
// headers and adds rounds only if particles are left (Gauss tables / scratch pool grow there).  Particles are dealt
// heaviest-first round-robin into groups on their own streams so that one group's dependent chain (tables -> assembly ->
// controller) overlaps the others' kernels.  The controller and the per-particle planner are plain functions (hs_dvctl.h)
// that the CPU test suite runs against the oracle's recorded trial sequences (tests/test_emu_cpu.py).
// Included by hydrotm.cu after hs_staged_host.h.
#pragma once
#include "hs_dvctl.h"

namespace {

using hs::DvP;

enum { DQ_ITEMS = 0, DQ_MODES, DQ_POOL, DQ_RAD, DQ_AT0, DQ_AT1, DQ_AS0, DQ_AS1, DQ_G3, DQ_G4, DQ_SM, DV_NQ };
enum { DK_RAD = 0, DK_WIG, DK_AT1, DK_AT0, DK_AS1, DK_AS0, DK_G4, DK_G3, DK_SM, DV_NK };

struct DvHdr {  // one group's round header
    long long tot[DV_NQ];   // list lengths / pool doubles of the planned round (after the capacity cut)
    hs::StQueue q[DV_NK];
    int active;             // unfinished particles counted by k_dv_plan
    int remaining;          // ... published by k_dv_scan
    int nmax_max;           // largest converged NMAX
    int arena_full;
    int gneed;              // largest Gauss rule an item asked for that the tables do not hold yet (0: none)
    int stuck;              // the first unfinished particle alone exceeds the list / pool capacity
    int rounds;             // rounds that had work
    int cut;                // first particle behind the capacity cut of the planned round (np: none)
    long long want_pool;    // pool doubles the round would have used without the cut (sizing of the next call)
    long long stuck_pool;   // stuck: pool doubles that particle needs
    int any_big, pad;       // some round had modes with NM > 32
};

struct DvCaps { long long cap[DV_NQ]; };

struct DvLists {
    hs::StItem *items;
    hs::StResult *res;
    int *modes;
    hs::StRadTask *rad;
    hs::StATask *at0, *at1, *as0, *as1;
    int *g3, *g4, *sm;
};

struct DvGroupBufs {
    DvP *P = nullptr; size_t p_cap = 0;
    long long *cnt = nullptr, *pre = nullptr, *btot = nullptr, *boff = nullptr; size_t cnt_cap = 0;  // [DV_NQ][np], [DV_NQ][nblocks]
    char *lists = nullptr; size_t lists_cap = 0;
    double *pool = nullptr; size_t pool_cap = 0;
    DvHdr *hdr = nullptr;
    cudaStream_t s1 = nullptr, s2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_done = nullptr;
};

struct DvShared {
    DvGroupBufs g[ST_GROUPS];
    int *d_n0 = nullptr, *d_lb = nullptr, *d_hist = nullptr; size_t n_cap = 0;
    DvHdr *h_hdr = nullptr;  // pinned copy of the groups' headers
    cudaEvent_t ev_in = nullptr;
    int occ[DV_NK] = {0};
    int last_rounds = 0;     // rounds the previous call needed (enqueued up front by the next one)
    double pool_per_particle = 0.0;  // doubles: grown from `want_pool` of earlier calls
    long long pool_floor = 0;        // doubles: largest single-particle need that once exceeded the pool
    bool had_big = true;             // the previous call had modes with NM > 32 (separate solve kernels at full grid)
    bool init = false;
};

__device__ __forceinline__ long long dv_item_pool(int nmax, int gp, int nmodes)
{
    // [tab | wig | sys] of one item, contiguous in the group's pool (sys: only modes with NM > 32 leave the SM)
    const int nbig = nmodes < nmax - 31 ? nmodes : (nmax - 31 > 0 ? nmax - 31 : 0);
    return hs::st_tab_size(nmax, gp) + hs::st_wig_mode_off(nmax, nmodes, gp) + 2 * hs::st_sys_mode_off(nmax, nbig);
}

// ---- setup: starting orders, heaviest-first deal into groups ----
__device__ __forceinline__ void dv_nmax0(int i, const double *axi, const double *lam, const double *mr, const double *mi, const double *eps,
                                         double rat, int np, int &start, int &lb)
{
    const double PI = 3.14159265358979323846;
    double r = rat;
    const double e_ = eps[i];
    if (fabs(rat - 1.0) > 1e-8 && np == -2) {  // SAREAC
        r = pow(1.5 / e_, 1.0 / 3.0);
        r = r / sqrt((e_ + 2.0) / (2.0 * e_));
    } else if (fabs(rat - 1.0) > 1e-8) {
        double e, q;
        if (e_ < 1.0) { e = sqrt(1.0 - e_ * e_); q = 0.5 * (pow(e_, 2.0 / 3.0) + pow(e_, -1.0 / 3.0) * asin(e) / e); }
        else { e = sqrt(1.0 - 1.0 / (e_ * e_)); q = 0.25 * (2.0 * pow(e_, 2.0 / 3.0) + pow(e_, -4.0 / 3.0) * log((1.0 + e) / (1.0 - e)) / e); }
        r = 1.0 / sqrt(q);
    }
    const double xev = 2.0 * PI * r * axi[i] / lam[i];
    const int ixxx = (xev > 0.0 && xev < 1e6) ? (int)(xev + 4.05 * pow(xev, 0.333333)) : 0;
    const int n0 = ixxx > 4 ? ixxx : 4;
    const double a = (e_ > 0.0) ? pow(e_, 1.0 / 3.0) : 1.0;
    double xa = xev * fmax(a, a / e_);
    if (np == -2 && e_ > 0.0) xa = xev * pow(2.0 / (3.0 * e_ * e_), 1.0 / 3.0) * sqrt(1.0 + e_ * e_);
    const double mx = sqrt(mr[i] * mr[i] + mi[i] * mi[i]) * xa;
    const bool bad = !(axi[i] > 0.0) || !(lam[i] > 0.0) || !(e_ > 0.0) || !isfinite(mr[i]) || !isfinite(mi[i]) || !(mr[i] * mr[i] + mi[i] * mi[i] > 0.0);
    start = bad ? -(n0 + 1) : n0;  // bad input: the starting order is still reported (as the fused kernel does)
    lb = (mx > 0.0 && mx < 1e6) ? max(n0 + 1, (int)floor(0.6035 * mx + 4.49) - 2) : n0 + 1;
}

__global__ void k_dv_hist(int n, const double *axi, const double *lam, const double *mr, const double *mi, const double *eps, double rat,
                          int np, int *n0, int *lb, int *hist)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int s, l;
    dv_nmax0(i, axi, lam, mr, mi, eps, rat, np, s, l);
    n0[i] = s; lb[i] = l;
    const int key = min(max(l, 0), HS_NPN1) + 1;  // deal by the predicted converged order
    atomicAdd(&hist[s < 0 ? 0 : key], 1);
}

struct DvGroupPtrs { DvP *P[ST_GROUPS]; };

__global__ void k_dv_place(int n, int ng, int spec0, const int *n0, const int *lb, int *hist, DvGroupPtrs G)
{
    // bin offsets, heaviest first (every block computes the same 103 numbers), then a slot inside the bin by atomics:
    // the order inside a bin is arbitrary (particles are independent; only their arena slots depend on it)
    __shared__ int base[HS_NPN1 + 3];
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int v = HS_NPN1 + 1; v >= 0; v--) { base[v] = acc; acc += hist[v]; }
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s0 = n0[i], l = lb[i];
    const int key = s0 < 0 ? 0 : min(max(l, 0), HS_NPN1) + 1;
    const int r = base[key] + atomicAdd(&hist[HS_NPN1 + 3 + key], 1);
    DvP s;
    hs::dv_init(s, i, s0, l, spec0);
    G.P[r % ng][r / ng] = s;
}

// ---- controller + count of the next round ----
struct DvPlanArgs {
    DvP *P; int np;
    const hs::StItem *items; const hs::StResult *res;
    long long *cnt;        // [DV_NQ][np] own counts
    long long *pre;        // [DV_NQ][np] exclusive prefix inside the particle's block of 256
    long long *btot;       // [DV_NQ][nblocks] block totals
    DvHdr *hdr;
    double ddelt; int ndgs, spec, margin, gmax, gfmax, shape;
    int *o_nmax, *o_ngauss, *o_ntrials, *o_status; long long *o_toff;
};

__global__ void __launch_bounds__(256) k_dv_plan(DvPlanArgs A)  // DV_PB threads
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    bool act = false;
    long long c[DV_NQ];
#pragma unroll
    for (int q = 0; q < DV_NQ; q++) c[q] = 0;
    if (j < A.np) {
    DvP s = A.P[j];
    // controller (hs_dvctl.h) over the results of the previous round, then the limit test of the coming one
    hs::dv_consume(s, A.ddelt, A.ndgs, A.margin, [&](int q) {
        const hs::StResult r = A.res[s.first_item + q];
        hs::DvTrial t;
        t.qsca = r.qsca; t.qext = r.qext; t.t_off = r.t_off; t.sing = r.sing; t.g = A.items[s.first_item + q].g;
        return t;
    });
    if (!s.done) {
        act = true;
        int gmiss = 0;
        const bool planned = hs::dv_items_of(s, A.ndgs, A.spec, A.gmax, A.gfmax, A.shape, &gmiss, [&](int nmax, int g, int nmodes) {
            const int gp = (g + 3) & ~3;
            c[DQ_ITEMS]++;
            c[DQ_MODES] += nmodes;
            c[DQ_POOL] += dv_item_pool(nmax, gp, nmodes);
            c[DQ_RAD] += (g + 15) / 16;
            for (int m = 0; m < nmodes; m++) {
                const int NM = nmax - (m > 1 ? m : 1) + 1;
                if (NM <= ST_SMALL_NM) { c[DQ_SM]++; continue; }
                if (NM <= 16) { c[m == 0 ? DQ_AS0 : DQ_AS1]++; continue; }
                const int nst = (NM + 31) / 32;
                c[m == 0 ? DQ_AT0 : DQ_AT1] += nst * nst;
                if (NM > 32) c[NM <= 56 ? DQ_G3 : DQ_G4]++;
            }
        });
        if (!planned) atomicMax(&A.hdr->gneed, gmiss);
    } else if (!s.written) {
        s.written = 1;
        int stt = s.status;
        if (stt == HS_ST_OK && s.sing) stt = HS_ST_SINGULAR;
        const int p = s.idx;
        A.o_nmax[p] = s.nmax; A.o_ngauss[p] = s.g; A.o_ntrials[p] = s.ntr; A.o_status[p] = stt;
        A.o_toff[p] = (s.done == 1) ? s.t_off : -1;
        if (s.done == 1) atomicMax(&A.hdr->nmax_max, s.nmax);
        if (s.status == HS_ST_ARENA_FULL) A.hdr->arena_full = 1;
    }
    A.P[j] = s;
#pragma unroll
    for (int q = 0; q < DV_NQ; q++) A.cnt[(size_t)q * A.np + j] = c[q];
    }
    const unsigned am = __ballot_sync(0xffffffffu, act);
    if ((threadIdx.x & 31) == 0 && am) atomicAdd(&A.hdr->active, __popc(am));
    // exclusive prefix of every count inside the block (k_dv_scan adds the block offsets)
    __shared__ long long wtot[DV_NQ][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long inc[DV_NQ];
#pragma unroll
    for (int q = 0; q < DV_NQ; q++) {
        long long v = c[q];
        for (int d = 1; d < 32; d <<= 1) { const long long u = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += u; }
        inc[q] = v;
        if (lane == 31) wtot[q][warp] = v;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < DV_NQ; q++) {
        long long base = 0, tot = 0;
        for (int w = 0; w < 8; w++) { const long long t = wtot[q][w]; if (w < warp) base += t; tot += t; }
        if (j < A.np) A.pre[(size_t)q * A.np + j] = base + inc[q] - c[q];
        if (threadIdx.x == 0) A.btot[(size_t)q * gridDim.x + blockIdx.x] = tot;
    }
}

// ---- list offsets: one CTA; warp q = exclusive prefix of the block totals of count q; capacity cut; queue lengths ----
#define DV_PB 256  // particles per k_dv_plan block
__global__ void __launch_bounds__(32 * DV_NQ) k_dv_scan(int np, int nb, const long long *cnt, const long long *pre, const long long *btot,
                                                          long long *boff, DvCaps caps, DvHdr *hdr, DvHdr *h_out)
{
    __shared__ long long s_tot[DV_NQ];
    __shared__ int s_bc, s_cut;
    const int t = threadIdx.x, lane = t & 31, q = t >> 5;
    if (t == 0) { s_bc = nb; s_cut = np; }
    {
        long long run = 0;
        for (int b0 = 0; b0 < nb; b0 += 32) {
            const int b = b0 + lane;
            const long long v = b < nb ? btot[(size_t)q * nb + b] : 0;
            long long inc = v;
            for (int d = 1; d < 32; d <<= 1) { const long long u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += u; }
            if (b < nb) boff[(size_t)q * nb + b] = run + inc - v;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) s_tot[q] = run;
    }
    __syncthreads();
    // capacity cut: the first particle whose lists / scratch would not fit, and everything after it, waits for the next
    // round (the inclusive prefixes are non-decreasing, so "does not fit" is monotone in the particle index)
    for (int b = t; b < nb; b += 32 * DV_NQ) {
        bool over = false;
        for (int k = 0; k < DV_NQ; k++) over |= boff[(size_t)k * nb + b] + btot[(size_t)k * nb + b] > caps.cap[k];
        if (over) { atomicMin(&s_bc, b); break; }
    }
    __syncthreads();
    const int bc = s_bc;
    if (bc < nb && t < DV_PB) {
        const int j = bc * DV_PB + t;
        if (j < np) {
            bool over = false;
            for (int k = 0; k < DV_NQ; k++) over |= boff[(size_t)k * nb + bc] + pre[(size_t)k * np + j] + cnt[(size_t)k * np + j] > caps.cap[k];
            if (over) atomicMin(&s_cut, j);
        }
    }
    __syncthreads();
    const int cut = s_cut;
    auto len = [&](int k) -> long long { return cut < np ? boff[(size_t)k * nb + bc] + pre[(size_t)k * np + cut] : s_tot[k]; };
    if (t < DV_NQ) {
        const long long tot = len(t);
        hdr->tot[t] = tot;
        if (t == DQ_POOL) hdr->want_pool = max(hdr->want_pool, s_tot[t]);
        if (t == DQ_ITEMS) {
            if (tot > 0) hdr->rounds++;
            if (cut < np && tot == 0) { hdr->stuck = 1; hdr->stuck_pool = cnt[(size_t)DQ_POOL * np + cut]; }
        }
    }
    if (t == 32) {
        hdr->q[DK_RAD].count = 3 * (int)len(DQ_RAD); hdr->q[DK_WIG].count = (int)len(DQ_MODES);
        hdr->q[DK_AT1].count = (int)len(DQ_AT1); hdr->q[DK_AT0].count = (int)len(DQ_AT0);
        hdr->q[DK_AS1].count = (int)len(DQ_AS1); hdr->q[DK_AS0].count = (int)len(DQ_AS0);
        hdr->q[DK_G4].count = (int)len(DQ_G4); hdr->q[DK_G3].count = (int)len(DQ_G3); hdr->q[DK_SM].count = (int)len(DQ_SM);
        if (len(DQ_G3) + len(DQ_G4) > 0) hdr->any_big = 1;
        for (int k = 0; k < DV_NK; k++) hdr->q[k].cursor = 0;
        hdr->remaining = hdr->active;
        hdr->active = 0;
        hdr->cut = cut;
    }
    // the host's copy of the header: written straight into pinned (mapped) host memory, so that the host only waits for
    // the stream -- a D2H copy of it would queue behind whatever bulk transfer the caller has in flight on the copy engine
    __syncthreads();
    if (t < (int)(sizeof(DvHdr) / sizeof(int))) reinterpret_cast<volatile int *>(h_out)[t] = reinterpret_cast<const int *>(hdr)[t];
}

// ---- lists of the round ----
struct DvFillArgs {
    DvP *P; int np;
    const long long *pre;   // [DV_NQ][np] prefix inside the block
    const long long *boff;  // [DV_NQ][nblocks] block offsets
    int nb;
    DvHdr *hdr;
    DvLists L;
    int ndgs, spec, gmax, gfmax, shape;
    long long arena_cap;
    unsigned long long *arena_top;
};

__global__ void __launch_bounds__(128) k_dv_fill(DvFillArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= A.np) return;
    if (j >= A.hdr->cut) return;  // behind the capacity cut: n_items stays 0
    DvP s = A.P[j];
    if (s.done) return;
    long long o[DV_NQ];
#pragma unroll
    for (int q = 0; q < DV_NQ; q++) o[q] = A.boff[(size_t)q * A.nb + j / DV_PB] + A.pre[(size_t)q * A.np + j];
    const long long tot_rad = A.hdr->tot[DQ_RAD];
    const int first = (int)o[DQ_ITEMS];
    int n_items = 0;
    const bool ok = hs::dv_items_of(s, A.ndgs, A.spec, A.gmax, A.gfmax, A.shape, nullptr, [&](int nmax, int g, int nmodes) {
        const int gp = (g + 3) & ~3;
        const int ii = (int)o[DQ_ITEMS]++;
        hs::StItem it;
        it.p = s.idx; it.nmax = nmax; it.g = g; it.gp = gp; it.nmodes = nmodes; it.first_mi = (int)o[DQ_MODES];
        it.tab_off = o[DQ_POOL];
        it.wig_off = it.tab_off + hs::st_tab_size(nmax, gp);
        it.sys_off = (it.wig_off + hs::st_wig_mode_off(nmax, nmodes, gp)) >> 1;  // complex units of the same pool
        o[DQ_POOL] += dv_item_pool(nmax, gp, nmodes);
        // final items: slot of the packed T in the caller's arena (a retried final trial keeps its slot)
        long long to = -1;
        if (nmodes > 1) {
            to = s.t_off;
            if (to < 0) {
                const unsigned long long sz = (unsigned long long)hs::tmat_size(nmax);
                const unsigned long long off = atomicAdd(A.arena_top, sz);
                to = (off + sz > (unsigned long long)A.arena_cap) ? -2 : (long long)off;
            }
        }
        it.t_off = to;
        A.L.items[ii] = it;
        hs::StResult r; r.qsca = 0.0; r.qext = 0.0; r.t_off = to; r.sing = 0; r.pad = 0;
        A.L.res[ii] = r;
        for (int kind = 0; kind < 3; kind++) {
            hs::StRadTask *R = A.L.rad + (2 - kind) * tot_rad + o[DQ_RAD];
            int k = 0;
            for (int node0 = 0; node0 < g; node0 += 16) R[k++] = {ii, (node0 << 2) | kind};
        }
        o[DQ_RAD] += (g + 15) / 16;
        for (int m = 0; m < nmodes; m++) {
            const int mix = hs::st_pack_mode(ii, m);
            A.L.modes[o[DQ_MODES]++] = mix;
            const int NM = nmax - (m > 1 ? m : 1) + 1;
            if (NM <= ST_SMALL_NM) { A.L.sm[o[DQ_SM]++] = mix; continue; }
            if (NM <= 16) { if (m == 0) A.L.as0[o[DQ_AS0]++] = {mix, 0, 0}; else A.L.as1[o[DQ_AS1]++] = {mix, 0, 0}; continue; }
            const int nst = (NM + 31) / 32;
            for (int sr = 0; sr < nst; sr++)
                for (int sc = 0; sc < nst; sc++) {
                    if (m == 0) A.L.at0[o[DQ_AT0]++] = {mix, (short)sr, (short)sc};
                    else A.L.at1[o[DQ_AT1]++] = {mix, (short)sr, (short)sc};
                }
            if (NM > 32) { if (NM <= 56) A.L.g3[o[DQ_G3]++] = mix; else A.L.g4[o[DQ_G4]++] = mix; }
        }
        n_items++;
    });
    if (!ok) return;
    A.P[j].first_item = first;
    A.P[j].n_items = n_items;
}

static int dv_init(hs_ctx *ctx, DvShared &D)
{
    if (D.init) return 0;
    CK(cudaEventCreateWithFlags(&D.ev_in, cudaEventDisableTiming));
    CK(cudaMallocHost((void **)&D.h_hdr, ST_GROUPS * sizeof(DvHdr)));
    for (int g = 0; g < ST_GROUPS; g++) {
        DvGroupBufs &b = D.g[g];
        CK(cudaStreamCreateWithFlags(&b.s1, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&b.s2, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&b.ev_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&b.ev_join, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&b.ev_done, cudaEventDisableTiming));
        CK(cudaMalloc((void **)&b.hdr, sizeof(DvHdr)));
    }
    using namespace hs;
    CK(cudaFuncSetAttribute(k_st_assemble_q<true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_ASM_SMEM));
    CK(cudaFuncSetAttribute(k_st_assemble_q<false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_ASM_SMEM));
    CK(cudaFuncSetAttribute(k_st_solve_q<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200704));
    CK(cudaFuncSetAttribute(k_st_mode_small_q, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * ST_SMALL_SLICE * 8));
    // resident CTAs per SM of every queue kernel
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&D.occ[DK_RAD], k_st_radial_q, 128, 0));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&D.occ[DK_WIG], k_st_wigner_q, 128, 0));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&D.occ[DK_AT1], k_st_assemble_q<false, 32>, 256, ST_ASM_SMEM));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&D.occ[DK_AT0], k_st_assemble_q<true, 32>, 256, ST_ASM_SMEM));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&D.occ[DK_AS1], k_st_assemble_q<false, 16>, 128, ST_ASM_SMEM16));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&D.occ[DK_AS0], k_st_assemble_q<true, 16>, 128, ST_ASM_SMEM16));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&D.occ[DK_G4], k_st_solve_q<256, true>, 256, 0));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&D.occ[DK_G3], k_st_solve_q<256, false>, 256, 200704));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&D.occ[DK_SM], k_st_mode_small_q, 256, 8 * ST_SMALL_SLICE * 8));
    for (int k = 0; k < DV_NK; k++) if (D.occ[k] < 1) D.occ[k] = 1;
    D.init = true;
    return 0;
}

static void dv_destroy(DvShared &D)
{
    if (!D.init) return;
    for (int g = 0; g < ST_GROUPS; g++) {
        DvGroupBufs &b = D.g[g];
        cudaFree(b.P); cudaFree(b.cnt); cudaFree(b.pre); cudaFree(b.btot); cudaFree(b.boff); cudaFree(b.lists); cudaFree(b.pool); cudaFree(b.hdr);
        if (b.s1) cudaStreamDestroy(b.s1);
        if (b.s2) cudaStreamDestroy(b.s2);
        if (b.ev_fork) cudaEventDestroy(b.ev_fork);
        if (b.ev_join) cudaEventDestroy(b.ev_join);
        if (b.ev_done) cudaEventDestroy(b.ev_done);
    }
    cudaFree(D.d_n0); cudaFree(D.d_lb); cudaFree(D.d_hist);
    if (D.h_hdr) cudaFreeHost(D.h_hdr);
    if (D.ev_in) cudaEventDestroy(D.ev_in);
    D.init = false;
}

}  // namespace

// Converged T-matrices of all n particles of the batch through device-driven rounds.  `st` is the caller's stream: the
// groups' streams wait for it (inputs) and it waits for them (outputs) before this returns.
static int calctmat_dev(hs_ctx *ctx, StShared &S, DvShared &D, cudaStream_t st, int n, const StCall &C)
{
    using namespace hs;
    if (st_init(ctx, S)) return -1;
    if (dv_init(ctx, D)) return -1;
    auto T0 = std::chrono::steady_clock::now();
    const int ng_auto = n >= 8192 ? 3 : (n >= 1024 ? 2 : 1);
    const int ng = std::max(1, std::min(ST_GROUPS, getenv("HYDROTM_DEV_GROUPS") ? atoi(getenv("HYDROTM_DEV_GROUPS")) : ng_auto));
    static const int spec0 = getenv("HYDROTM_SPEC0") ? atoi(getenv("HYDROTM_SPEC0")) : 0;
    static const int margin = getenv("HYDROTM_SPEC_MARGIN") ? atoi(getenv("HYDROTM_SPEC_MARGIN")) : 1;
    const int sms = ctx->sm_count;

    // ---- buffers ----
    if ((size_t)n > D.n_cap) {
        cudaFree(D.d_n0); cudaFree(D.d_lb); cudaFree(D.d_hist);
        D.d_n0 = D.d_lb = D.d_hist = nullptr; D.n_cap = 0;
        CK(cudaMalloc((void **)&D.d_n0, (size_t)n * sizeof(int)));
        CK(cudaMalloc((void **)&D.d_lb, (size_t)n * sizeof(int)));
        CK(cudaMalloc((void **)&D.d_hist, 2 * (HS_NPN1 + 3) * sizeof(int)));
        D.n_cap = n;
    }
    const int npg_max = (n + ng - 1) / ng;
    DvCaps caps;
    DvLists L[ST_GROUPS];
    {
        const long long ci = std::min<long long>((long long)(12 + std::max(spec0, 0) + std::max(C.spec, 0)) * npg_max + 64, 1ll << 23);
        const long long cm = 48ll * npg_max + 8192;
        caps.cap[DQ_ITEMS] = ci; caps.cap[DQ_MODES] = cm; caps.cap[DQ_RAD] = 6 * ci + 4096;
        caps.cap[DQ_AT0] = cm; caps.cap[DQ_AT1] = cm; caps.cap[DQ_AS0] = cm; caps.cap[DQ_AS1] = cm;
        caps.cap[DQ_G3] = cm; caps.cap[DQ_G4] = cm; caps.cap[DQ_SM] = cm;
        // pool: what earlier calls asked for per particle (+25 %), at least one worst-case item, at most the budget share
        const double per = D.pool_per_particle > 0.0 ? 1.25 * D.pool_per_particle : 12288.0;  // doubles
        long long pool = (long long)(per * npg_max);
        pool = std::max<long long>(pool, 16ll << 20);                                       // 128 MB: one NMAX 100 / NGAUSS 500 item
        pool = std::min<long long>(pool, std::max<long long>((long long)(C.scratch_budget / 8 / ng), 1ll << 17));
        pool = std::max<long long>(pool, D.pool_floor);  // a particle of an earlier call needed this much on its own
        caps.cap[DQ_POOL] = pool & ~1ll;
    }
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    for (int g = 0; g < ng; g++) {
        DvGroupBufs &b = D.g[g];
        if (st_grow(ctx, &b.P, &b.p_cap, (size_t)npg_max)) return -1;
        if ((size_t)npg_max * DV_NQ > b.cnt_cap) {
            cudaFree(b.cnt); cudaFree(b.pre); cudaFree(b.btot); cudaFree(b.boff); b.cnt = b.pre = b.btot = b.boff = nullptr; b.cnt_cap = 0;
            CK(cudaMalloc((void **)&b.cnt, (size_t)npg_max * DV_NQ * sizeof(long long)));
            CK(cudaMalloc((void **)&b.pre, (size_t)npg_max * DV_NQ * sizeof(long long)));
            CK(cudaMalloc((void **)&b.btot, (size_t)(npg_max / DV_PB + 1) * DV_NQ * sizeof(long long)));
            CK(cudaMalloc((void **)&b.boff, (size_t)(npg_max / DV_PB + 1) * DV_NQ * sizeof(long long)));
            b.cnt_cap = (size_t)npg_max * DV_NQ;
        }
        const size_t ci = (size_t)caps.cap[DQ_ITEMS], cm = (size_t)caps.cap[DQ_MODES], cr = (size_t)caps.cap[DQ_RAD];
        const size_t o_items = 0, o_res = al(o_items + ci * sizeof(StItem)), o_mi = al(o_res + ci * sizeof(StResult)),
                     o_rad = al(o_mi + cm * 4), o_a0 = al(o_rad + 3 * cr * sizeof(StRadTask)), o_a1 = al(o_a0 + cm * sizeof(StATask)),
                     o_s0 = al(o_a1 + cm * sizeof(StATask)), o_s1 = al(o_s0 + cm * sizeof(StATask)), o_g3 = al(o_s1 + cm * sizeof(StATask)),
                     o_g4 = al(o_g3 + cm * 4), o_sm = al(o_g4 + cm * 4), o_end = al(o_sm + cm * 4);
        if (st_grow(ctx, &b.lists, &b.lists_cap, o_end)) return -1;
        if (st_grow(ctx, &b.pool, &b.pool_cap, (size_t)caps.cap[DQ_POOL] + 2)) return -1;
        char *d = b.lists;
        L[g].items = (StItem *)(d + o_items); L[g].res = (StResult *)(d + o_res); L[g].modes = (int *)(d + o_mi);
        L[g].rad = (StRadTask *)(d + o_rad); L[g].at0 = (StATask *)(d + o_a0); L[g].at1 = (StATask *)(d + o_a1);
        L[g].as0 = (StATask *)(d + o_s0); L[g].as1 = (StATask *)(d + o_s1); L[g].g3 = (int *)(d + o_g3); L[g].g4 = (int *)(d + o_g4);
        L[g].sm = (int *)(d + o_sm);
    }
    if (ensure_gauss(ctx, 128)) return -1;
    if (C.np == -2 && ensure_gauss_full(ctx, 65)) return -1;

    // ---- setup on the caller's stream ----
    CK(cudaMemsetAsync(D.d_hist, 0, 2 * (HS_NPN1 + 3) * sizeof(int), st));
    for (int g = 0; g < ng; g++) CK(cudaMemsetAsync(D.g[g].hdr, 0, sizeof(DvHdr), st));
    k_dv_hist<<<(n + 255) / 256, 256, 0, st>>>(n, C.d_axi, C.d_lam, C.d_mr, C.d_mi, C.d_eps, C.rat, C.np, D.d_n0, D.d_lb, D.d_hist);
    DvGroupPtrs GP;
    for (int g = 0; g < ST_GROUPS; g++) GP.P[g] = D.g[g].P;
    k_dv_place<<<(n + 255) / 256, 256, 0, st>>>(n, ng, spec0, D.d_n0, D.d_lb, D.d_hist, GP);
    long long launches = 2;
    CK(cudaGetLastError());
    CK(cudaEventRecord(D.ev_in, st));

    int npg[ST_GROUPS];
    for (int g = 0; g < ng; g++) { npg[g] = (n - g + ng - 1) / ng; CK(cudaStreamWaitEvent(D.g[g].s1, D.ev_in, 0)); }

    auto stage_args = [&](int g) {
        StArgs a;
        a.items = L[g].items; a.modes = L[g].modes;
        a.axi = C.d_axi; a.lam = C.d_lam; a.mr = C.d_mr; a.mi = C.d_mi; a.eps = C.d_eps; a.rat = C.rat;
        a.gx = ctx->d_gx; a.gw = ctx->d_gw; a.gfx = ctx->d_gfx; a.gfw = ctx->d_gfw; a.np = C.np; a.cm = S.d_cm; a.dd = S.d_dd;
        a.tab = D.g[g].pool; a.wig = D.g[g].pool; a.sys = reinterpret_cast<double2 *>(D.g[g].pool);
        a.tarena = C.d_tarena; a.arena_cap = C.arena_cap; a.arena_top = C.d_arena_top; a.toff = C.d_toff; a.res = L[g].res;
        a.fuse = 1 | 2 | 4 | ST_FUSE_RESERVED;
        if (getenv("HYDROTM_SOLVE_REG")) { const int v = atoi(getenv("HYDROTM_SOLVE_REG")); a.fuse = 1 | 2 | (v == 0 ? 0 : (v == 2 ? 12 : 4)) | ST_FUSE_RESERVED; }
        if (!(getenv("HYDROTM_SOLVE_TILE") && atoi(getenv("HYDROTM_SOLVE_TILE")) == 0)) a.fuse |= ST_FUSE_TILE;
        a.prof = C.prof;
        return a;
    };
    auto plan = [&](int g) -> int {
        DvGroupBufs &b = D.g[g];
        DvPlanArgs A;
        A.P = b.P; A.np = npg[g]; A.items = L[g].items; A.res = L[g].res; A.cnt = b.cnt; A.pre = b.pre; A.btot = b.btot; A.hdr = b.hdr;
        const int nb = (npg[g] + DV_PB - 1) / DV_PB;
        A.ddelt = C.ddelt; A.ndgs = C.ndgs; A.spec = C.spec; A.margin = margin; A.gmax = ctx->gmax; A.gfmax = ctx->gfmax; A.shape = C.np;
        A.o_nmax = C.d_nmax; A.o_ngauss = C.d_ngauss; A.o_ntrials = C.d_ntrials; A.o_status = C.d_status; A.o_toff = C.d_toff;
        k_dv_plan<<<nb, DV_PB, 0, b.s1>>>(A);
        k_dv_scan<<<1, 32 * DV_NQ, 0, b.s1>>>(npg[g], nb, b.cnt, b.pre, b.btot, b.boff, caps, b.hdr, &D.h_hdr[g]);
        launches += 2;
        CK(cudaGetLastError());
        return 0;
    };
    const bool one_lane = C.one_lane;
    auto round = [&](int g) -> int {
        DvGroupBufs &b = D.g[g];
        DvFillArgs F;
        F.P = b.P; F.np = npg[g]; F.pre = b.pre; F.boff = b.boff; F.nb = (npg[g] + DV_PB - 1) / DV_PB; F.hdr = b.hdr; F.L = L[g];
        F.ndgs = C.ndgs; F.spec = C.spec; F.gmax = ctx->gmax; F.gfmax = ctx->gfmax; F.shape = C.np;
        F.arena_cap = C.arena_cap; F.arena_top = C.d_arena_top;
        cudaStream_t s1 = b.s1, s2 = one_lane ? b.s1 : b.s2;
        k_dv_fill<<<(npg[g] + 127) / 128, 128, 0, s1>>>(F);
        const StArgs a = stage_args(g);
        static const int psel = getenv("HYDROTM_PROF_SEL") ? atoi(getenv("HYDROTM_PROF_SEL")) : 31;  // see hs_staged_host.h
        auto pa = [&](int bit) { StArgs x = a; if (!(psel & bit)) x.prof = nullptr; return x; };
        StQueue *q = b.hdr->q;
        // resident CTAs per SM of the radial / Wigner kernels (HYDROTM_RAD_CTAS, HYDROTM_WIG_CTAS: experiments)
        static const int rad_ctas = getenv("HYDROTM_RAD_CTAS") ? std::max(1, atoi(getenv("HYDROTM_RAD_CTAS"))) : 8;
        static const int wig_ctas = getenv("HYDROTM_WIG_CTAS") ? std::max(1, atoi(getenv("HYDROTM_WIG_CTAS"))) : 8;
        k_st_radial_q<<<sms * std::min(D.occ[DK_RAD], rad_ctas), 128, 0, s1>>>(a, L[g].rad, q + DK_RAD);
        k_st_wigner_q<<<sms * std::min(D.occ[DK_WIG], wig_ctas), 128, 0, s1>>>(a, q + DK_WIG);
        if (!one_lane) { CK(cudaEventRecord(b.ev_fork, s1)); CK(cudaStreamWaitEvent(s2, b.ev_fork, 0)); }
        k_st_assemble_q<false, 32><<<sms * D.occ[DK_AT1], 256, ST_ASM_SMEM, s1>>>(pa(1), L[g].at1, q + DK_AT1);
        k_st_assemble_q<true, 32><<<sms * D.occ[DK_AT0], 256, ST_ASM_SMEM, s1>>>(pa(2), L[g].at0, q + DK_AT0);
        // modes with NM > 32 (hail-sized particles).  When the previous call had none, a quarter of the grid is launched: the
        // empty 200 KB-per-CTA launches cost ~1 % of a rain-sweep call at full size, and a call that does bring such modes
        // after one that did not still drains its lists at a quarter of the rate instead of crawling on a token grid
        k_st_solve_q<256, true><<<D.had_big ? sms * std::min(D.occ[DK_G4], 2) : sms / 2, 256, 0, s1>>>(a, L[g].g4, q + DK_G4);
        k_st_solve_q<256, false><<<D.had_big ? sms * D.occ[DK_G3] : sms / 4, 256, 200704, s1>>>(a, L[g].g3, q + DK_G3);
        k_st_assemble_q<false, 16><<<sms * D.occ[DK_AS1], 128, ST_ASM_SMEM16, s2>>>(pa(4), L[g].as1, q + DK_AS1);
        k_st_assemble_q<true, 16><<<sms * D.occ[DK_AS0], 128, ST_ASM_SMEM16, s2>>>(pa(8), L[g].as0, q + DK_AS0);
        k_st_mode_small_q<<<sms * D.occ[DK_SM], 256, 8 * ST_SMALL_SLICE * 8, s2>>>(pa(16), L[g].sm, q + DK_SM);
        if (!one_lane) { CK(cudaEventRecord(b.ev_join, s2)); CK(cudaStreamWaitEvent(s1, b.ev_join, 0)); }
        launches += 10;
        CK(cudaGetLastError());
        return 0;
    };
    auto read_hdrs = [&]() -> int {
        for (int g = 0; g < ng; g++) CK(cudaStreamSynchronize(D.g[g].s1));  // k_dv_scan has written D.h_hdr[g]
        return 0;
    };

    // HYDROTM_DEV_TIMELINE=1: an event after every round of every group (GPU time since the call's inputs were ready)
    const bool timeline = getenv("HYDROTM_DEV_TIMELINE") != nullptr;
    struct TL { cudaEvent_t ev; int g, r; };
    std::vector<TL> tl;
    cudaEvent_t ev_t0 = nullptr;
    if (timeline) { cudaEventCreate(&ev_t0); cudaEventRecord(ev_t0, st); for (int g = 0; g < ng; g++) cudaStreamWaitEvent(D.g[g].s1, ev_t0, 0); }
    auto mark = [&](int g, int r) { if (!timeline) return; TL e; e.g = g; e.r = r; cudaEventCreate(&e.ev); cudaEventRecord(e.ev, D.g[g].s1); tl.push_back(e); };
    int rounds_done = 0;
    int batch = D.last_rounds > 0 ? D.last_rounds : 6;
    if (getenv("HYDROTM_DEV_ROUNDS")) batch = std::max(1, atoi(getenv("HYDROTM_DEV_ROUNDS")));
    int rc = 0;
    for (int g = 0; g < ng && !rc; g++) rc = plan(g);
    for (; !rc;) {
        for (int r = 0; r < batch && !rc; r++)
            for (int g = 0; g < ng && !rc; g++) { rc = round(g); if (!rc) rc = plan(g); mark(g, rounds_done + r); }
        rounds_done += batch;
        if (rc) break;
        if (read_hdrs()) { rc = -1; break; }
        int remaining = 0, gneed = 0, stuck = 0;
        for (int g = 0; g < ng; g++) { remaining += D.h_hdr[g].remaining; gneed = std::max(gneed, D.h_hdr[g].gneed); stuck |= D.h_hdr[g].stuck; }
        if (C.trace) fprintf(stderr, "[hydrotm-dev] %d rounds enqueued: %d particles left, gneed %d\n", rounds_done, remaining, gneed);
        if (!remaining) break;
        const bool grow_gauss = gneed > ctx->gmax;
        if (stuck) {
            // a particle that does not fit the pool on its own (tiny scratch budget, or NMAX / NGAUSS near the limits):
            // the pool grows to what it needs -- every group stream is idle here (read_hdrs synchronised them)
            long long need = 0;
            for (int g = 0; g < ng; g++) if (D.h_hdr[g].stuck) need = std::max(need, D.h_hdr[g].stuck_pool);
            if (need <= caps.cap[DQ_POOL]) { hs_set_err(ctx, "device-driven rounds: a particle exceeds the task-list capacity"); rc = -4; break; }
            need = (need + 1) & ~1ll;
            D.pool_floor = std::max(D.pool_floor, need);
            caps.cap[DQ_POOL] = need;
            bool fail = false;
            for (int g = 0; g < ng && !fail; g++) fail = st_grow(ctx, &D.g[g].pool, &D.g[g].pool_cap, (size_t)need + 2) != 0;
            if (fail) { rc = -1; break; }
        }
        if (grow_gauss) {
            if (ensure_gauss(ctx, std::min(HS_NPNG1, std::max(gneed, 2 * ctx->gmax)))) { rc = -1; break; }
            if (C.np == -2 && ensure_gauss_full(ctx, ctx->gmax / 2 + 1)) { rc = -1; break; }
        }
        if (stuck || grow_gauss) {
            // the round already planned was counted against the old tables / pool: plan again (the controller part of
            // k_dv_plan is idempotent: the results of the last round have been consumed)
            bool fail = false;
            for (int g = 0; g < ng && !fail; g++) {
                fail = cudaMemsetAsync(&D.g[g].hdr->gneed, 0, sizeof(int), D.g[g].s1) != cudaSuccess ||
                       cudaMemsetAsync(&D.g[g].hdr->stuck, 0, sizeof(int), D.g[g].s1) != cudaSuccess || plan(g) != 0;
            }
            if (fail) { hs_set_err(ctx, "device-driven rounds: re-planning failed"); rc = -1; break; }
        }
        if (rounds_done > 4 * HS_NPN1 + HS_NPNG1) { hs_set_err(ctx, "device-driven controller did not terminate"); rc = -4; break; }
        batch = 2;
    }
    // join the group streams back into the caller's stream (also on errors: the caller may drop its buffers)
    for (int g = 0; g < ng; g++) {
        cudaError_t e = cudaEventRecord(D.g[g].ev_done, D.g[g].s1);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, D.g[g].ev_done, 0);
        if (e != cudaSuccess && !rc) { hs_set_err(ctx, cudaGetErrorString(e)); rc = -1; }
    }
    ctx->launches += launches;
    if (timeline) {
        cudaDeviceSynchronize();
        for (auto &e : tl) { float ms = -1.f; cudaEventElapsedTime(&ms, ev_t0, e.ev); cudaEventDestroy(e.ev); fprintf(stderr, "[dev-timeline] g%d r%d done %7.3f ms\n", e.g, e.r, ms); }
        cudaEventDestroy(ev_t0);
    }
    if (rc) { cudaDeviceSynchronize(); return rc; }
    int need_rounds = 0;
    double want = 0.0;
    for (int g = 0; g < ng; g++) {
        const DvHdr &h = D.h_hdr[g];
        if (h.nmax_max > C.nmax_max) C.nmax_max = h.nmax_max;
        if (h.arena_full) C.arena_full = true;
        need_rounds = std::max(need_rounds, h.rounds);
        want = std::max(want, (double)h.want_pool / (double)std::max(1, npg[g]));
    }
    D.last_rounds = std::max(1, need_rounds);
    D.had_big = false;
    for (int g = 0; g < ng; g++) D.had_big |= D.h_hdr[g].any_big != 0;
    if (want > D.pool_per_particle) D.pool_per_particle = want;
    if (C.trace) {
        const double tot = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T0).count();
        fprintf(stderr, "[hydrotm-dev] %d particles, %d groups, %d rounds with work (%d enqueued), pool %.1f MB per group (wanted %.1f), %.2f ms\n",
                n, ng, need_rounds, rounds_done, (double)caps.cap[DQ_POOL] * 8 / 1048576.0, want * npg_max * 8 / 1048576.0, tot);
    }
    return 0;
}
