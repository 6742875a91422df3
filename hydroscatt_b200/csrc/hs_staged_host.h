// hs_staged_host.h -- host side of the staged T-matrix pipeline (hs_staged.cuh): the NMAX / NGAUSS convergence
// controller of `calctmat` run between rounds of stage kernels.
//
// Round = { plan items -> one H2D of the lists -> reserve, S, W, A, G kernels -> one D2H of (Qsca, Qext) -> controller }.
// Trials are independent computations, so each round runs K speculative trials per unconverged particle
// (NMAX, NMAX+1, ...; the controller consumes them in the reference's order and discards the ones past convergence)
// and, once NMAX has converged, the first NGAUSS trial together with every mode m >= 1 of that configuration
// ("final" item): on the rain sweep the NGAUSS phase always converges at its first trial, so the T-matrix is
// complete in the same round.
// The particles are dealt into ST_GROUPS independent groups, each with its own stream, scratch and round loop; the
// host plans / consumes one group while the others' kernels run, so neither the host work nor the tail of a stage
// kernel leaves the GPU idle.  Included by hydrotm.cu after hs_ctx / CK are defined.
#pragma once

namespace {

#define ST_GROUPS 8

struct StBufs {
    char *d_lists = nullptr; size_t lists_cap = 0;    // items | mitems | rad tasks | A tasks (m=0) | A tasks | G lists | small list
    char *h_lists = nullptr; size_t h_lists_cap = 0;  // pinned staging of the same
    hs::StResult *h_res = nullptr; size_t res_cap = 0;  // pinned + mapped: the stage kernels write their results straight into it
    double *d_tab = nullptr; size_t tab_cap = 0;
    double *d_wig = nullptr; size_t wig_cap = 0;
    double2 *d_sys = nullptr; size_t sys_cap = 0;
    int *d_out = nullptr, *h_out = nullptr; size_t out_cap = 0;  // final per-particle records for the scatter-out kernel
    cudaStream_t stream = nullptr;
    cudaEvent_t ev = nullptr;
    cudaStream_t stream2 = nullptr;   // second lane of a round: the narrow / small-mode kernels run beside the wide ones
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // fused scatter: records / order list of finished particles, filled round by round
    cudaStream_t sc_stream = nullptr;
    cudaEvent_t sc_ev = nullptr;
    int *d_screc = nullptr, *h_screc = nullptr; size_t d_screc_cap = 0, h_screc_cap = 0;  // 8 ints per particle: idx nmax g ntr status - toff(lo,hi)
    int *d_scord = nullptr, *h_scord = nullptr; size_t d_scord_cap = 0, h_scord_cap = 0;
};

// host-side planning vectors of a group, kept between calls so that steady-state calls do not reallocate them
struct StHostVecs {
    std::vector<hs::StItem> items;
    std::vector<int> mitems, gl[5], sm, order;
    std::vector<hs::StRadTask> rad;
    std::vector<hs::StATask> at0, at1, as0, as1;
};

// host threads that drive the groups, kept between calls (a call used to spawn and join its own: VERDICT r1 item 7)
struct StPool {
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv, cv_done;
    std::function<void(int)> job;
    unsigned long long gen = 0;
    int pending = 0, nth = 0;
    bool stop = false;
    void worker(int id)
    {
        unsigned long long seen = 0;
        for (;;) {
            std::function<void(int)> j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || (gen != seen && id < nth); });
                if (stop) return;
                seen = gen;
                j = job;
            }
            j(id);
            {
                std::lock_guard<std::mutex> lk(mu);
                if (--pending == 0) cv_done.notify_all();
            }
        }
    }
    // run f(0) .. f(n - 1) on n pool threads and wait for all of them
    void run(int n, std::function<void(int)> f)
    {
        while ((int)th.size() < n) { const int id = (int)th.size(); th.emplace_back([this, id] { worker(id); }); }
        {
            std::lock_guard<std::mutex> lk(mu);
            job = std::move(f); nth = n; pending = n; gen++;
        }
        cv.notify_all();
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return pending == 0; });
    }
    ~StPool()
    {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        for (auto &t : th) t.join();
    }
};

struct StShared {
    StBufs g[ST_GROUPS];
    StHostVecs hv[ST_GROUPS];
    StPool pool;
    double *d_cm = nullptr, *d_dd = nullptr;
    cudaEvent_t ev_in = nullptr;
};

struct StP {  // controller state of one particle (Ctl of the fused kernel)
    int idx, phase, nmax, g, ntr, status, done, sing;
    double q1s, q1e;
    long long t_off;
    int first_item, n_items;  // this round's items
    int scattered;            // fused scatter: record written / scatter kernel launched
    int kspec;                // speculative trials to run next round
    double dprev, dlast;      // max(dsca, dext) / ddelt of the last two consumed trials (convergence trend)
};

template <class T>
static int st_grow(hs_ctx *ctx, T **p, size_t *cap, size_t need, bool pinned = false)
{
    if (need <= *cap) return 0;
    size_t ncap = std::max(need, *cap + *cap / 2);
    if (*p) { if (pinned) cudaFreeHost(*p); else cudaFree(*p); }
    *p = nullptr; *cap = 0;
    if (pinned) CK(cudaMallocHost((void **)p, ncap * sizeof(T)));
    else CK(cudaMalloc((void **)p, ncap * sizeof(T)));
    *cap = ncap;
    return 0;
}

__global__ void k_st_scatter_out(int n, const int *rec, const long long *toffs, int *nmax, int *ngauss, int *ntrials,
                                 int *status, long long *toff)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = rec[5 * i];
    nmax[p] = rec[5 * i + 1];
    ngauss[p] = rec[5 * i + 2];
    ntrials[p] = rec[5 * i + 3];
    status[p] = rec[5 * i + 4];
    toff[p] = toffs[i];
}

__global__ void k_st_scatter_out8(int n, const int *rec, int *nmax, int *ngauss, int *ntrials, int *status, long long *toff)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int *r = rec + 8 * (long)i;
    const int p = r[0];
    nmax[p] = r[1];
    ngauss[p] = r[2];
    ntrials[p] = r[3];
    status[p] = r[4];
    toff[p] = (long long)(((unsigned long long)(unsigned)r[7] << 32) | (unsigned)r[6]);
}

struct StCall {  // arguments of one hs_calctmat_batch call
    const double *d_axi, *d_lam, *d_mr, *d_mi, *d_eps;
    double rat, ddelt;
    int np = -1;  // shape (HS_SHAPE_*)
    int ndgs;
    double *d_tarena;
    long long arena_cap;
    unsigned long long *d_arena_top;
    long long *d_toff;
    int *d_nmax, *d_ngauss, *d_ntrials, *d_status;
    size_t scratch_budget;
    int spec;
    bool trace;
    bool fuse_solve = true;
    bool one_lane = false;           // HYDROTM_ONE_LANE=1: every kernel of a round on one stream
    bool kr16 = true;                // modes with NM <= 16 through the 128-thread variant of k_st_assemble
    unsigned long long *prof = nullptr;
    const ScPrep *sp = nullptr;      // fused scatter plan (hs_calctmat_scatter_batch) or null
    mutable int nmax_max = 0;        // out: largest converged NMAX among the staged particles
    mutable bool arena_full = false;  // out: some particle did not fit the caller's arena
};

struct StGroup {
    hs_ctx *ctx;
    StShared *S;
    StBufs *B;
    const StCall *C;
    int id = 0;
    std::vector<StP> P;
    std::vector<hs::StItem> items;
    std::vector<int> mitems;  // packed (item, mode)
    std::vector<hs::StRadTask> rad;
    std::vector<hs::StATask> at0, at1, as0, as1;  // A tasks: KR = 32 kernel (m = 0, m >= 1), KR = 16 kernel (NM <= 16)
    std::vector<int> gl[5], sm;
    std::vector<int> order;
    int nround = 0, K = 2;
    bool inflight = false;
    double t_plan = 0, t_wait = 0, t_ctl = 0;
    long long launches = 0;
    int rc = 0;
    struct TL { cudaEvent_t ev; const char *what; int round; double host_ms; };
    std::vector<TL> tl;  // HYDROTM_TIMELINE=1: an event after every launch
    bool timeline = false;
    std::chrono::steady_clock::time_point t_base;
    void mark(const char *what)
    {
        if (!timeline) return;
        TL e; e.what = what; e.round = nround;
        e.host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_base).count();
        cudaEventCreate(&e.ev); cudaEventRecord(e.ev, B->stream); tl.push_back(e);
    }

    void swap_vecs(StHostVecs &h)
    {
        items.swap(h.items); mitems.swap(h.mitems); sm.swap(h.sm); order.swap(h.order); rad.swap(h.rad);
        at0.swap(h.at0); at1.swap(h.at1); as0.swap(h.as0); as1.swap(h.as1);
        for (int c = 0; c < 5; c++) gl[c].swap(h.gl[c]);
    }

    bool active() const
    {
        for (const StP &s : P) if (!s.done) return true;
        return false;
    }

    // plan the next round and enqueue it; returns 0 (inflight may stay false if nothing was left to launch)
    int issue()
    {
        using namespace hs;
        auto T0 = std::chrono::steady_clock::now();
        const int ndgs = C->ndgs;
        cudaStream_t st = B->stream;
        order.clear();
        for (int j = 0; j < (int)P.size(); j++) if (!P[j].done) order.push_back(j);
        if (order.empty()) return 0;
        {   // heaviest particles first (counting sort on the current order)
            std::vector<int> cnt(HS_NPN1 + 2, 0), tmp(order.size());
            for (int j : order) cnt[std::min(std::max(P[j].nmax, 0), HS_NPN1)]++;
            int acc = 0;
            for (int v = HS_NPN1; v >= 0; v--) { int c = cnt[v]; cnt[v] = acc; acc += c; }
            for (int j : order) tmp[cnt[std::min(std::max(P[j].nmax, 0), HS_NPN1)]++] = j;
            order.swap(tmp);
        }
        const int active_n = (int)order.size();
        K = C->spec > 0 ? C->spec : (active_n > 1000 ? 2 : (active_n > 200 ? 3 : 4));
        items.clear(); mitems.clear(); rad.clear(); at0.clear(); at1.clear(); as0.clear(); as1.clear(); sm.clear();
        for (auto &v : gl) v.clear();
        long long tab_off = 0, wig_off = 0, sys_off = 0;
        int gneed = 0;
        for (int j : order) {
            StP &s = P[j];
            s.first_item = (int)items.size(); s.n_items = 0;
            if (s.phase == 0 && (long long)s.nmax * ndgs > HS_NPNG1) { s.status = HS_ST_NGAUSS_LIMIT; s.done = 2; continue; }
            // next round: scratch budget spent, or the packed (item, mode) ids (24-bit item index) would overflow
            if (((size_t)(tab_off + wig_off) * 8 + (size_t)sys_off * 16 > C->scratch_budget || items.size() > (1u << 23)) && !items.empty()) continue;
            const int cnt = s.phase == 0 ? (C->spec > 0 ? K : s.kspec) : 1;
            for (int q = 0; q < cnt; q++) {
                StItem it;
                it.p = s.idx;
                it.nmax = s.phase == 0 ? s.nmax + q : s.nmax;
                it.g = s.phase == 0 ? it.nmax * ndgs : (s.phase == 1 ? s.g + 1 : s.g);
                if (it.nmax > HS_NPN1 || it.g > HS_NPNG1) break;  // the controller reports the limit when it gets there
                it.gp = (it.g + 3) & ~3;
                it.nmodes = s.phase == 0 ? 1 : it.nmax + 1;
                it.first_mi = (int)mitems.size();
                it.tab_off = tab_off; it.wig_off = wig_off; it.sys_off = sys_off;
                it.t_off = s.phase == 0 ? -1 : s.t_off;
                tab_off += st_tab_size(it.nmax, it.gp);
                wig_off += st_wig_mode_off(it.nmax, it.nmodes, it.gp);
                sys_off += C->fuse_solve ? st_sys_mode_off(it.nmax, std::min(it.nmodes, std::max(0, it.nmax - 31)))  // only modes with NM > 32 leave the SM
                                         : st_sys_mode_off(it.nmax, it.nmodes);
                gneed = std::max(gneed, it.g);
                const int ii = (int)items.size();
                for (int m = 0; m < it.nmodes; m++) {
                    const int mix = st_pack_mode(ii, m);
                    mitems.push_back(mix);
                    const int NM = it.nmax - (m > 1 ? m : 1) + 1;
                    if (NM <= ST_SMALL_NM) { sm.push_back(mix); continue; }
                    if (NM <= 16 && C->kr16) { (m == 0 ? as0 : as1).push_back({mix, 0, 0}); if (!C->fuse_solve) gl[0].push_back(mix); continue; }
                    const int nst = (NM + 31) / 32;
                    for (int sr = 0; sr < nst; sr++)
                        for (int sc = 0; sc < nst; sc++) (m == 0 ? at0 : at1).push_back({mix, (short)sr, (short)sc});
                    if (NM > 32) gl[NM <= 56 ? 3 : 4].push_back(mix);  // smaller modes are solved inside k_st_assemble
                    else if (!C->fuse_solve) gl[NM <= 16 ? 0 : (NM <= 24 ? 1 : 2)].push_back(mix);
                }
                items.push_back(it);
                s.n_items++;
            }
        }
        if (items.empty()) return 0;
        // radial tasks of 16 nodes, kind-major (the complex-argument recurrence, the longest, first)
        for (int kind = 2; kind >= 0; kind--)
            for (int ii = 0; ii < (int)items.size(); ii++)
                for (int node0 = 0; node0 < items[ii].g; node0 += 16) rad.push_back({ii, (node0 << 2) | kind});
        if (gneed > ctx->gmax) {
            std::lock_guard<std::mutex> lk(ctx->mu);
            if (ensure_gauss(ctx, std::max(gneed, 64))) return -1;
        }
        if (C->np == -2 && gneed / 2 + 1 > ctx->gfmax) {
            std::lock_guard<std::mutex> lk(ctx->mu);
            if (ensure_gauss_full(ctx, gneed / 2 + 1)) return -1;
        }
        const double *gx, *gw, *gfx, *gfw;
        { std::lock_guard<std::mutex> lk(ctx->mu); gx = ctx->d_gx; gw = ctx->d_gw; gfx = ctx->d_gfx; gfw = ctx->d_gfw; }
        auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
        const size_t o_items = 0, o_mi = al(o_items + items.size() * sizeof(StItem)), o_rad = al(o_mi + mitems.size() * sizeof(int)),
                     o_a0 = al(o_rad + rad.size() * sizeof(StRadTask)), o_a1 = al(o_a0 + at0.size() * sizeof(StATask)),
                     o_s0 = al(o_a1 + at1.size() * sizeof(StATask)), o_s1 = al(o_s0 + as0.size() * sizeof(StATask)),
                     o_g0 = al(o_s1 + as1.size() * sizeof(StATask)), o_g1 = al(o_g0 + gl[0].size() * 4), o_g2 = al(o_g1 + gl[1].size() * 4),
                     o_g3 = al(o_g2 + gl[2].size() * 4), o_g4 = al(o_g3 + gl[3].size() * 4), o_sm = al(o_g4 + gl[4].size() * 4), o_end = al(o_sm + sm.size() * 4);
        if (st_grow(ctx, &B->d_lists, &B->lists_cap, o_end)) return -1;
        if (st_grow(ctx, &B->h_lists, &B->h_lists_cap, o_end, true)) return -1;
        if (st_grow(ctx, &B->h_res, &B->res_cap, items.size(), true)) return -1;
        if (st_grow(ctx, &B->d_tab, &B->tab_cap, (size_t)tab_off)) return -1;
        if (st_grow(ctx, &B->d_wig, &B->wig_cap, (size_t)wig_off)) return -1;
        if (st_grow(ctx, &B->d_sys, &B->sys_cap, (size_t)std::max<long long>(sys_off, 1))) return -1;
        char *h = B->h_lists, *d = B->d_lists;
        memcpy(h + o_items, items.data(), items.size() * sizeof(StItem));
        memcpy(h + o_mi, mitems.data(), mitems.size() * sizeof(int));
        memcpy(h + o_rad, rad.data(), rad.size() * sizeof(StRadTask));
        memcpy(h + o_a0, at0.data(), at0.size() * sizeof(StATask));
        memcpy(h + o_a1, at1.data(), at1.size() * sizeof(StATask));
        memcpy(h + o_s0, as0.data(), as0.size() * sizeof(StATask));
        memcpy(h + o_s1, as1.data(), as1.size() * sizeof(StATask));
        const size_t og[6] = {o_g0, o_g1, o_g2, o_g3, o_g4, o_sm};
        for (int c = 0; c < 6; c++) {
            // heaviest systems first (counting sort on NM), written straight into the staging buffer
            const std::vector<int> &L = c < 5 ? gl[c] : sm;
            int cnt[HS_NPN1 + 2] = {0};
            for (int x : L) cnt[items[x & 0xFFFFFF].nmax - std::max(x >> 24, 1) + 1]++;
            int acc = 0;
            for (int v = HS_NPN1; v >= 0; v--) { int q = cnt[v]; cnt[v] = acc; acc += q; }
            int *dst = (int *)(h + og[c]);
            for (int x : L) dst[cnt[items[x & 0xFFFFFF].nmax - std::max(x >> 24, 1) + 1]++] = x;
        }
        mark("begin");
        CK(cudaMemcpyAsync(d, h, o_end, cudaMemcpyHostToDevice, st));
        mark("h2d");
        StArgs a;
        a.items = (const StItem *)(d + o_items); a.modes = (const int *)(d + o_mi);
        a.axi = C->d_axi; a.lam = C->d_lam; a.mr = C->d_mr; a.mi = C->d_mi; a.eps = C->d_eps; a.rat = C->rat;
        a.gx = gx; a.gw = gw; a.gfx = gfx; a.gfw = gfw; a.np = C->np; a.cm = S->d_cm; a.dd = S->d_dd;
        a.tab = B->d_tab; a.wig = B->d_wig; a.sys = B->d_sys;
        a.tarena = C->d_tarena; a.arena_cap = C->arena_cap; a.arena_top = C->d_arena_top; a.toff = C->d_toff; a.res = B->h_res; a.fuse = ((getenv("HYDROTM_SOLVE_TILE") && atoi(getenv("HYDROTM_SOLVE_TILE")) == 0) ? 0 : ST_FUSE_TILE) | (C->fuse_solve ? 1 : 0) | (getenv("HYDROTM_SOLVE2") && atoi(getenv("HYDROTM_SOLVE2")) == 0 ? 0 : 2) | (getenv("HYDROTM_SOLVE_REG") ? (atoi(getenv("HYDROTM_SOLVE_REG")) == 0 ? 0 : (atoi(getenv("HYDROTM_SOLVE_REG")) == 2 ? 12 : 4)) : 4); a.prof = C->prof;
        const int ni = (int)items.size();
        // HYDROTM_PROF_SEL: bit mask of the kernel variants that record phase cycles (1: m>=1 wide, 2: m=0 wide, 4: m>=1 narrow,
        // 8: m=0 narrow, 16: small modes); default all
        static const int psel = getenv("HYDROTM_PROF_SEL") ? atoi(getenv("HYDROTM_PROF_SEL")) : 31;
        auto pa = [&](int bit) { StArgs x = a; if (!(psel & bit)) x.prof = nullptr; return x; };
        k_st_radial<<<((int)rad.size() + 7) / 8, 128, 0, st>>>(a, (const StRadTask *)(d + o_rad), (int)rad.size());
        mark("radial");
        k_st_wigner<<<((int)mitems.size() + 7) / 8, 128, 0, st>>>(a, (int)mitems.size());
        mark("wigner");
        launches += 2;
        // two lanes after the tables are ready: the wide kernels (NM > 16, multi-block, separate solves) on the group's
        // stream, the narrow (NM <= 16) and small-mode (NM <= 8) kernels on its second stream
        cudaStream_t s2 = st;
        const bool two = !C->one_lane && (!at1.empty() || !at0.empty()) && (!as1.empty() || !as0.empty() || !sm.empty());
        if (two) {
            s2 = B->stream2;
            CK(cudaEventRecord(B->ev_fork, st));
            CK(cudaStreamWaitEvent(s2, B->ev_fork, 0));
        }
        if (!at1.empty()) { k_st_assemble<false, 32><<<(int)at1.size(), 256, ST_ASM_SMEM, st>>>(pa(1), (const StATask *)(d + o_a1), (int)at1.size()); launches++; mark("asm m>=1"); }
        if (!at0.empty()) { k_st_assemble<true, 32><<<(int)at0.size(), 256, ST_ASM_SMEM, st>>>(pa(2), (const StATask *)(d + o_a0), (int)at0.size()); launches++; mark("asm m=0"); }
        if (!gl[4].empty()) { k_st_solve<256, true><<<(int)gl[4].size(), 256, 0, st>>>(a, (const int *)(d + o_g4), (int)gl[4].size()); launches++; }
        const int gsm[4] = {16384, 36864, 65536, 200704};
        if (!gl[3].empty()) { k_st_solve<256, false><<<(int)gl[3].size(), 256, gsm[3], st>>>(a, (const int *)(d + og[3]), (int)gl[3].size()); launches++; }
        for (int c = 2; c >= (C->kr16 ? 1 : 0); c--)
            if (!gl[c].empty()) { k_st_solve<256, false><<<(int)gl[c].size(), 256, gsm[c], st>>>(a, (const int *)(d + og[c]), (int)gl[c].size()); launches++; mark("solve"); }
        if (!as1.empty()) { k_st_assemble<false, 16><<<(int)as1.size(), 128, ST_ASM_SMEM16, s2>>>(pa(4), (const StATask *)(d + o_s1), (int)as1.size()); launches++; }
        if (!as0.empty()) { k_st_assemble<true, 16><<<(int)as0.size(), 128, ST_ASM_SMEM16, s2>>>(pa(8), (const StATask *)(d + o_s0), (int)as0.size()); launches++; }
        if (C->kr16 && !gl[0].empty()) {
            // right-sized solver for NM <= 16 (64 threads, shared memory = the mode's two systems); the list is sorted by NM (descending):
            // two launches so that the small modes are not held to the shared-memory footprint of NM = 16
            const int *L = (const int *)(h + og[0]);
            const int tot = (int)gl[0].size();
            auto nm_of = [&](int x) { return items[x & 0xFFFFFF].nmax - std::max(x >> 24, 1) + 1; };
            int cut = 0;
            while (cut < tot && nm_of(L[cut]) > 11) cut++;
            if (cut) { const int nm = nm_of(L[0]); k_st_solve<64, false><<<cut, 64, 64 * nm * nm, s2>>>(a, (const int *)(d + og[0]), cut); launches++; }
            if (tot > cut) { const int nm = nm_of(L[cut]); k_st_solve<64, false><<<tot - cut, 64, 64 * nm * nm, s2>>>(a, (const int *)(d + og[0]) + cut, tot - cut); launches++; }
        }
        if (!sm.empty()) { k_st_mode_small<<<((int)sm.size() + 7) / 8, 256, 8 * ST_SMALL_SLICE * 8, s2>>>(pa(16), (const int *)(d + o_sm), (int)sm.size()); launches++; }
        if (two) {
            CK(cudaEventRecord(B->ev_join, s2));
            CK(cudaStreamWaitEvent(st, B->ev_join, 0));
        }
        mark("round end");
        CK(cudaGetLastError());
        inflight = true;
        if (C->trace)
            fprintf(stderr, "[hydrotm-staged] group %d round %d: %d active, K=%d, %d items, %zu mode-items (%zu small), %zu+%zu (+%zu+%zu narrow) A tasks, scratch %.1f MB\n",
                    id, nround, active_n, K, ni, mitems.size(), sm.size(), at0.size(), at1.size(), as0.size(), as1.size(),
                    ((double)(tab_off + wig_off) * 8 + (double)sys_off * 16) / 1048576.0);
        t_plan += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T0).count();
        return 0;
    }

    // wait for the round in flight and run the controller (calctmat_particle: the block after HS_PROF(3))
    int finish()
    {
        using namespace hs;
        if (!inflight) return 0;
        auto T0 = std::chrono::steady_clock::now();
        CK(cudaStreamSynchronize(B->stream));
        auto T1 = std::chrono::steady_clock::now();
        if (timeline) { TL e; e.ev = nullptr; e.what = "host: synced"; e.round = nround; e.host_ms = std::chrono::duration<double, std::milli>(T1 - t_base).count(); tl.push_back(e); }
        inflight = false;
        const double ddelt = C->ddelt;
        for (int j : order) {
            StP &s = P[j];
            if (s.done || s.n_items == 0) continue;  // limit hit while planning, or deferred by the scratch budget
            for (int q = 0; q < s.n_items; q++) {
                const StItem &it = items[s.first_item + q];
                const StResult &r = B->h_res[s.first_item + q];
                if (s.phase == 2) {  // modes of a configuration whose convergence is already decided
                    if (r.t_off == -2) { s.status = HS_ST_ARENA_FULL; s.done = 2; break; }
                    s.t_off = r.t_off; s.sing |= r.sing; s.done = 1;
                    break;
                }
                const double dsca = fabs((s.q1s - r.qsca) / r.qsca), dext = fabs((s.q1e - r.qext) / r.qext);
                s.q1s = r.qsca; s.q1e = r.qext;
                s.ntr++;
                s.g = it.g;
                s.sing |= r.sing;
                const bool ok = (dsca <= ddelt && dext <= ddelt);
                if (s.phase == 0) { s.dprev = s.dlast; s.dlast = std::max(dsca, dext) / ddelt; }
                if (s.phase == 0) {
                    if (ok) { s.phase = (it.g == HS_NPNG1) ? 2 : 1; break; }
                    else if (s.nmax + 1 > HS_NPN1) { s.status = HS_ST_NMAX_LIMIT; s.done = 2; break; }
                    else s.nmax = s.nmax + 1;
                } else {
                    if (r.t_off == -2) { s.status = HS_ST_ARENA_FULL; s.done = 2; break; }
                    s.t_off = r.t_off;
                    if (ok) s.done = 1;
                    else if (it.g == HS_NPNG1) { s.status = HS_ST_NGAUSS_LIMIT; s.done = 1; }
                    break;
                }
            }
        }
        // speculation depth of the next round from the convergence trend: the relative change of Qsca / Qext shrinks roughly
        // geometrically with NMAX, so log(d_last) / log(d_prev / d_last) more trials bring it under DDELT
        for (int j : order) {
            StP &s = P[j];
            if (s.done || s.phase != 0) continue;
            // HYDROTM_SPEC_MARGIN (default 1) trials beyond the predicted one; without a trend (a single relative change so
            // far) one trial if the change is already within 100 x DDELT, else two.  Under-speculation costs a round for that
            // particle, over-speculation costs discarded work: the pipeline is throughput-bound, but a shorter chain of rounds measured slightly faster than the leanest speculation (margin 0: 6 rounds, 7.6 ms; 1: 5 rounds, 7.4 ms).
            static const int margin = getenv("HYDROTM_SPEC_MARGIN") ? atoi(getenv("HYDROTM_SPEC_MARGIN")) : 1;
            int k = (s.dlast > 0.0 && s.dlast < 100.0) ? 1 : 2;
            if (s.dprev > s.dlast && s.dlast > 1.0 && std::isfinite(s.dprev)) {
                const double r = log(s.dlast) / log(s.dprev / s.dlast);
                k = (int)std::min(8.0, std::max(1.0, ceil(r) + (double)margin));
            }
            s.kspec = k;
        }
        nround++;
        if (nround > 4 * HS_NPN1 + HS_NPNG1) { hs_set_err(ctx, "staged controller did not terminate"); return -4; }
        auto T2 = std::chrono::steady_clock::now();
        t_wait += std::chrono::duration<double, std::milli>(T1 - T0).count();
        t_ctl += std::chrono::duration<double, std::milli>(T2 - T1).count();
        return 0;
    }

    // Fused scatter (hs_calctmat_scatter_batch): particles that reached a terminal state since the last call get their
    // per-particle record written and their orientation-averaged S / Z / cross sections computed right away on the
    // group's scatter stream, while the remaining particles go through further convergence rounds.
    int sc_filled = 0;
    int flush_scatter()
    {
        using namespace hs;
        const ScPrep *sp = C->sp;
        if (!sp) return 0;
        const int ns = (int)P.size();
        if (st_grow(ctx, &B->h_screc, &B->h_screc_cap, (size_t)ns * 8, true)) return -1;
        if (st_grow(ctx, &B->d_screc, &B->d_screc_cap, (size_t)ns * 8)) return -1;
        if (st_grow(ctx, &B->h_scord, &B->h_scord_cap, (size_t)ns, true)) return -1;
        if (st_grow(ctx, &B->d_scord, &B->d_scord_cap, (size_t)ns)) return -1;
        // newly finished particles, heaviest first
        int cnt[HS_NPN1 + 2] = {0};
        int nnew = 0, bound = 1;
        for (const StP &s : P)
            if (s.done && !s.scattered) { cnt[s.done == 1 ? std::min(std::max(s.nmax, 0), HS_NPN1) : 0]++; nnew++; }
        if (!nnew) return 0;
        int acc = 0;
        for (int v = HS_NPN1; v >= 0; v--) { int c = cnt[v]; if (c && v > bound) bound = v; cnt[v] = acc; acc += c; }
        const int off = sc_filled;
        for (StP &s : P) {
            if (!s.done || s.scattered) continue;
            s.scattered = 1;
            const int q = off + cnt[s.done == 1 ? std::min(std::max(s.nmax, 0), HS_NPN1) : 0]++;
            int *r = B->h_screc + 8 * (size_t)q;
            int stt = s.status;
            if (stt == HS_ST_OK && s.sing) stt = HS_ST_SINGULAR;
            const long long to = (s.done == 1) ? s.t_off : -1;
            r[0] = s.idx; r[1] = s.nmax; r[2] = s.g; r[3] = s.ntr; r[4] = stt; r[5] = 0;
            r[6] = (int)(unsigned)((unsigned long long)to & 0xFFFFFFFFull); r[7] = (int)(unsigned)((unsigned long long)to >> 32);
            B->h_scord[q] = s.idx;
        }
        sc_filled += nnew;
        cudaStream_t ss = B->sc_stream;
        CK(cudaMemcpyAsync(B->d_screc + 8 * (size_t)off, B->h_screc + 8 * (size_t)off, (size_t)nnew * 8 * sizeof(int), cudaMemcpyHostToDevice, ss));
        CK(cudaMemcpyAsync(B->d_scord + off, B->h_scord + off, (size_t)nnew * sizeof(int), cudaMemcpyHostToDevice, ss));
        k_st_scatter_out8<<<(nnew + 255) / 256, 256, 0, ss>>>(nnew, B->d_screc + 8 * (size_t)off, C->d_nmax, C->d_ngauss, C->d_ntrials, C->d_status, C->d_toff);
        const size_t smem = sc_smem(sp->threads, bound);
        if (smem > (size_t)ctx->smem_optin - 8192) { hs_set_err(ctx, "scatter: nmax too large for the shared-memory T stage"); return -4; }
        ScArgs sa = sp->a;
        sa.order = B->d_scord + off;
        sa.nmax_bound = bound;
        CK(sc_launch(1, nnew, sp->threads, smem, ss, sa));
        launches += 2;
        CK(cudaGetLastError());
        return 0;
    }

    // the group's whole life on its own host thread
    void run()
    {
        if (cudaSetDevice(ctx->device) != cudaSuccess) { rc = -1; return; }
        rc = issue();
        while (!rc && inflight) advance();
        finalize();
    }

    // one controller turn: the round in flight is (or is waited to be) complete -> controller, next round, scatter
    void advance()
    {
        rc = finish();
        if (!rc) rc = issue();       // keep this group's next round ahead of its scatter work in the launch queue
        if (!rc) rc = flush_scatter();
    }
    void finalize()
    {
        if (!rc) rc = C->sp ? flush_scatter() : write_out();
    }
    bool round_ready() const { return cudaStreamQuery(B->stream) == cudaSuccess; }

    // Several groups on ONE host thread (fewer host cores than groups, e.g. 8 ranks on a 32-core box): whichever group's
    // round has completed gets its controller turn; the others keep the GPU busy meanwhile.
    static void run_many(StGroup *G, int first, int step, int ng)
    {
        if (first + step >= ng) { G[first].run(); return; }
        if (cudaSetDevice(G[first].ctx->device) != cudaSuccess) { for (int g = first; g < ng; g += step) G[g].rc = -1; return; }
        for (int g = first; g < ng; g += step) G[g].rc = G[g].issue();
        for (;;) {
            int live = 0, last = -1;
            bool progressed = false;
            for (int g = first; g < ng; g += step) {
                if (G[g].rc || !G[g].inflight) continue;
                live++; last = g;
                if (G[g].round_ready()) { G[g].advance(); progressed = true; }
            }
            if (!live) break;
            if (live == 1 && !progressed) G[last].advance();  // nothing else to look after: block on its stream
            else if (!progressed) std::this_thread::yield();
        }
        for (int g = first; g < ng; g += step) G[g].finalize();
    }

    // per-particle records -> device arrays
    int write_out()
    {
        const int ns = (int)P.size();
        if (!ns) return 0;
        cudaStream_t st = B->stream;
        if ((size_t)ns > B->out_cap) {
            if (B->d_out) cudaFree(B->d_out);
            if (B->h_out) cudaFreeHost(B->h_out);
            B->d_out = nullptr; B->h_out = nullptr; B->out_cap = 0;
            CK(cudaMalloc(&B->d_out, ((size_t)ns * 7 + 2) * sizeof(int)));
            CK(cudaMallocHost(&B->h_out, ((size_t)ns * 7 + 2) * sizeof(int)));
            B->out_cap = ns;
        }
        const size_t toffs_off = ((size_t)5 * ns + ((5 * ns) & 1));
        long long *h_toffs = reinterpret_cast<long long *>(B->h_out + toffs_off);
        for (int j = 0; j < ns; j++) {
            const StP &s = P[j];
            int stt = s.status;
            if (stt == HS_ST_OK && s.sing) stt = HS_ST_SINGULAR;
            B->h_out[5 * j] = s.idx; B->h_out[5 * j + 1] = s.nmax; B->h_out[5 * j + 2] = s.g; B->h_out[5 * j + 3] = s.ntr; B->h_out[5 * j + 4] = stt;
            h_toffs[j] = (s.done == 1) ? s.t_off : -1;
        }
        CK(cudaMemcpyAsync(B->d_out, B->h_out, ((size_t)ns * 7 + 2) * sizeof(int), cudaMemcpyHostToDevice, st));
        k_st_scatter_out<<<(ns + 255) / 256, 256, 0, st>>>(ns, B->d_out, reinterpret_cast<const long long *>(B->d_out + toffs_off), C->d_nmax,
                                                           C->d_ngauss, C->d_ntrials, C->d_status, C->d_toff);
        launches++;
        CK(cudaGetLastError());
        return 0;
    }
};

}  // namespace

static int st_init(hs_ctx *ctx, StShared &S)
{
    using namespace hs;
    if (S.d_cm) return 0;
    std::vector<double> cm(HS_NPN1 + 1), dd(HS_NPN1 + 1);
    double cacc = 1.0;
    cm[0] = 1.0; dd[0] = 0.0;
    for (int i = 1; i <= HS_NPN1; i++) {
        cacc = cacc * sqrt((double)(2 * i - 1) / (double)(2 * i));
        cm[i] = cacc;
        dd[i] = sqrt((double)(2 * i + 1) / (double)(i * (i + 1)));
    }
    CK(cudaMalloc(&S.d_cm, cm.size() * sizeof(double)));
    CK(cudaMalloc(&S.d_dd, dd.size() * sizeof(double)));
    CK(cudaMemcpy(S.d_cm, cm.data(), cm.size() * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(S.d_dd, dd.data(), dd.size() * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(k_st_assemble<true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_ASM_SMEM));
    CK(cudaFuncSetAttribute(k_st_assemble<false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_ASM_SMEM));
    CK(cudaFuncSetAttribute(k_st_solve<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200704));
    CK(cudaFuncSetAttribute(k_st_solve<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200704));
    CK(cudaFuncSetAttribute(k_st_mode_small, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * ST_SMALL_SLICE * 8));
    CK(cudaEventCreate(&S.ev_in));
    int prio_lo = 0, prio_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));  // opt-in scatter overlap: the scatter streams run at low priority
    for (int g = 0; g < ST_GROUPS; g++) {
        CK(cudaStreamCreateWithFlags(&S.g[g].stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&S.g[g].ev, cudaEventDisableTiming));
        CK(cudaStreamCreateWithPriority(&S.g[g].sc_stream, cudaStreamNonBlocking, prio_lo));
        CK(cudaStreamCreateWithFlags(&S.g[g].stream2, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&S.g[g].ev_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&S.g[g].ev_join, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&S.g[g].sc_ev, cudaEventDisableTiming));
    }
    return 0;
}

static void st_destroy(StShared &S)
{
    for (int g = 0; g < ST_GROUPS; g++) {
        StBufs &b = S.g[g];
        cudaFree(b.d_lists); cudaFree(b.d_tab); cudaFree(b.d_wig); cudaFree(b.d_sys); cudaFree(b.d_out);
        if (b.h_lists) cudaFreeHost(b.h_lists);
        if (b.h_res) cudaFreeHost(b.h_res);
        if (b.h_out) cudaFreeHost(b.h_out);
        if (b.stream) cudaStreamDestroy(b.stream);
        if (b.ev) cudaEventDestroy(b.ev);
        cudaFree(b.d_screc); cudaFree(b.d_scord);
        if (b.h_screc) cudaFreeHost(b.h_screc);
        if (b.h_scord) cudaFreeHost(b.h_scord);
        if (b.sc_stream) cudaStreamDestroy(b.sc_stream);
        if (b.stream2) cudaStreamDestroy(b.stream2);
        if (b.ev_fork) cudaEventDestroy(b.ev_fork);
        if (b.ev_join) cudaEventDestroy(b.ev_join);
        if (b.sc_ev) cudaEventDestroy(b.sc_ev);
    }
    cudaFree(S.d_cm); cudaFree(S.d_dd);
    if (S.ev_in) cudaEventDestroy(S.ev_in);
}

// Converged T-matrices of the particles `sel` (indices into the batch; n0_of = their starting NMAX, -1 = invalid input)
// through the staged pipeline.  `st` is the caller's stream: the groups' streams wait for it (inputs) and it waits for
// them (outputs) before this returns.
static int calctmat_staged(hs_ctx *ctx, StShared &S, cudaStream_t st, const std::vector<int> &sel, const std::vector<int> &n0_of,
                           const std::vector<int> &lb_of, const StCall &C)
{
    const int ns = (int)sel.size();
    if (!ns) return 0;
    if (st_init(ctx, S)) return -1;
    CK(cudaEventRecord(S.ev_in, st));
    // Groups (concurrent rounds on the GPU) and the host threads that drive them: one thread per group when the host has
    // the cores, else several groups per thread (leave a core for the caller's thread, and share the host between the
    // ranks of a one-process-per-GPU job: torchrun exports LOCAL_WORLD_SIZE)
    const int ng_auto = ns >= 8192 ? 6 : (ns >= 2048 ? 4 : (ns >= 256 ? 2 : 1));
    const int ng = std::max(1, std::min(ST_GROUPS, getenv("HYDROTM_GROUPS") ? atoi(getenv("HYDROTM_GROUPS")) : ng_auto));
    int nth = ng;
    {
        const int cores = (int)std::thread::hardware_concurrency();
        const int ranks = getenv("LOCAL_WORLD_SIZE") ? std::max(1, atoi(getenv("LOCAL_WORLD_SIZE"))) : 1;
        if (cores > 0) nth = std::max(1, std::min(ng, cores / ranks - 1));
        if (getenv("HYDROTM_THREADS")) nth = std::max(1, std::min(ng, atoi(getenv("HYDROTM_THREADS"))));
    }
    // deal the particles, heaviest first, round-robin: every group gets the same size mix
    // (stable counting sort on the starting order: a comparison sort of 17 220 keys cost 0.3 ms before the first kernel)
    std::vector<int> ord(ns);
    {
        int cnt[HS_NPN1 + 3] = {0};
        auto key = [&](int j) { return std::min(std::max(n0_of[j], -1), HS_NPN1) + 1; };
        for (int j = 0; j < ns; j++) cnt[key(j)]++;
        int acc = 0;
        for (int v = HS_NPN1 + 1; v >= 0; v--) { const int c = cnt[v]; cnt[v] = acc; acc += c; }
        for (int j = 0; j < ns; j++) ord[cnt[key(j)]++] = j;
    }
    StGroup G[ST_GROUPS];
    for (int g = 0; g < ng; g++) {
        G[g].ctx = ctx; G[g].S = &S; G[g].B = &S.g[g]; G[g].C = &C; G[g].id = g;
        G[g].swap_vecs(S.hv[g]);  // capacity of the previous call
        CK(cudaStreamWaitEvent(S.g[g].stream, S.ev_in, 0));
        CK(cudaStreamWaitEvent(S.g[g].sc_stream, S.ev_in, 0));
    }
    for (int g = 0; g < ng; g++) G[g].P.reserve((size_t)ns / ng + 1);
    for (int r = 0; r < ns; r++) {
        const int j = ord[r];
        StP s;
        s.idx = sel[j]; s.phase = 0; s.nmax = n0_of[j]; s.g = 0; s.ntr = 0; s.status = HS_ST_OK; s.done = 0; s.sing = 0;
        s.q1s = 0.0; s.q1e = 0.0; s.t_off = -1; s.first_item = 0; s.n_items = 0; s.scattered = 0;
        // every trial from the starting order to the converged one is needed, so a lower bound of the converged order is free
        // parallelism: the first round runs them all at once
        static const int spec0 = getenv("HYDROTM_SPEC0") ? atoi(getenv("HYDROTM_SPEC0")) : 0;  // extra speculative trials in round 0
        s.kspec = std::max(2, std::min(12 + spec0, lb_of[j] - n0_of[j] + 1 + spec0)); s.dprev = 0.0; s.dlast = 0.0;
        if (s.nmax < 0) { s.done = 2; s.status = HS_ST_BAD_INPUT; s.nmax = 0; }
        else if (s.nmax >= HS_NPN1) { s.done = 2; s.status = HS_ST_NMAX_LIMIT; }
        G[r % ng].P.push_back(s);
    }
    auto T0 = std::chrono::steady_clock::now();
    const bool timeline = getenv("HYDROTM_TIMELINE") != nullptr;
    for (int g = 0; g < ng; g++) { G[g].timeline = timeline; G[g].t_base = T0; }
    if (nth == 1) StGroup::run_many(G, 0, 1, ng);
    else S.pool.run(nth, [&G, nth, ng](int t) { StGroup::run_many(G, t, nth, ng); });
    for (int g = 0; g < ng; g++) G[g].swap_vecs(S.hv[g]);
    for (int g = 0; g < ng; g++)
        for (const StP &s : G[g].P) {
            if (s.done == 1 && s.nmax > C.nmax_max) C.nmax_max = s.nmax;
            if (s.status == HS_ST_ARENA_FULL) C.arena_full = true;
        }
    // join EVERY group stream back into the caller's stream before reporting an error: the caller may drop its arena and
    // output buffers as soon as this returns non-zero, and kernels of the healthy groups may still be running
    int first_rc = 0;
    for (int g = 0; g < ng; g++) {
        ctx->launches += G[g].launches;
        if (G[g].rc && !first_rc) first_rc = G[g].rc;
        cudaError_t e1 = cudaEventRecord(S.g[g].ev, S.g[g].stream);
        if (e1 == cudaSuccess) e1 = cudaStreamWaitEvent(st, S.g[g].ev, 0);
        cudaError_t e2 = cudaEventRecord(S.g[g].sc_ev, S.g[g].sc_stream);
        if (e2 == cudaSuccess) e2 = cudaStreamWaitEvent(st, S.g[g].sc_ev, 0);
        if ((e1 != cudaSuccess || e2 != cudaSuccess) && !first_rc) {
            hs_set_err(ctx, cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
            first_rc = -1;
        }
    }
    if (first_rc) { cudaDeviceSynchronize(); return first_rc; }
    if (timeline) {
        cudaDeviceSynchronize();
        for (int g = 0; g < ng; g++)
            for (auto &e : G[g].tl) {
                float ms = -1.f;
                if (e.ev) { cudaEventElapsedTime(&ms, S.ev_in, e.ev); cudaEventDestroy(e.ev); }
                fprintf(stderr, "[timeline] g%d r%d %-12s host %7.3f ms  gpu-done %7.3f ms\n", g, e.round, e.what, e.host_ms, ms);
            }
    }
    if (C.trace) {
        double tot = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T0).count();
        for (int g = 0; g < ng; g++)
            fprintf(stderr, "[hydrotm-staged] group %d: %zu particles, %d rounds: plan+issue %.2f ms, wait %.2f ms, controller %.2f ms\n", g,
                    G[g].P.size(), G[g].nround, G[g].t_plan, G[g].t_wait, G[g].t_ctl);
        fprintf(stderr, "[hydrotm-staged] %d particles, %d groups on %d host threads, %.2f ms\n", ns, ng, nth, tot);
    }
    return 0;
}
