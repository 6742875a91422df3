// hydrotm.cu -- C ABI of libhydrotm.so (see include/hydrotm.h) and the host-side bucketing / launch logic.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <chrono>
#include <string>
#include <vector>
#include <algorithm>
#include <mutex>
#include <thread>
#include <condition_variable>
#include <functional>

#include "../../include/hydrotm.h"
#include "hs_gauss.h"
#include "hs_tmat.cuh"
#include "hs_ampl.cuh"
#include "hs_xsect.cuh"
#include "hs_scatter.cuh"
#include "hs_psd.cuh"
#include "hs_cant.cuh"
#include "hs_cantmc.cuh"
#include "hs_tm2l.cuh"
#include "hs_staged.cuh"

struct hs_ctx {
    int device = 0;
    int sm_count = 148;
    int smem_optin = 227 * 1024;
    std::string err;
    long long launches = 0;
    int last_nmax_max = 0;       // largest converged NMAX seen by hs_calctmat_batch (sizes k_scatter's T stage)
    // Gauss table
    int gmax = 0;
    double *d_gx = nullptr, *d_gw = nullptr;
    int gfmax = 0;               // cylinders: full n-point Gauss rules, n = 1..gfmax
    double *d_gfx = nullptr, *d_gfw = nullptr;
    // scratch
    int *d_queue = nullptr;      // work-queue counters (16)
    int *d_items = nullptr;      // particle index lists
    int items_cap = 0;
    int *h_pin = nullptr;        // pinned staging
    size_t h_pin_cap = 0;
    cudaStream_t bstream[8] = {};  // one stream per size bucket
    cudaEvent_t ev_start = nullptr, ev_done[8] = {};
    double *d_sc = nullptr;      // scatter geometry tables (device)
    double *d_frames = nullptr;  // scatter kernel: per (orientation, geometry) frames
    size_t frames_cap = 0;
    double2 *d_scwig = nullptr;  // scatter kernel: Wigner table of the call (k_scatter_wig)
    size_t scwig_cap = 0;
    int *d_order = nullptr;      // scatter kernel: CTA -> particle map, heaviest first
    int order_cap = 0;
    double *d_xs_tmp = nullptr;
    size_t xs_tmp_cap = 0;
    long long *d_prof = nullptr; // phase-cycle counters (HYDROTM_PROF=1)
    double *d_ws_fb = nullptr;   // global-memory fall-back workspace slices
    double *d_ws_global = nullptr;
    size_t ws_global_bytes = 0;
    std::mutex mu;               // guards the Gauss table when the staged groups run on their own host threads
    std::mutex err_mu;           // guards err (written from the group threads; separate: CK runs inside regions that hold mu)
    std::vector<double *> gauss_old;  // superseded Gauss tables (kernels in flight may still read them; freed with the context)
    struct StBufsHolder *st = nullptr;  // staged pipeline buffers (hs_staged_host.h)
};

static inline void hs_set_err(hs_ctx *ctx, const std::string &msg)
{
    std::lock_guard<std::mutex> lk(ctx->err_mu);
    ctx->err = msg;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            hs_set_err(ctx, std::string(#call) + ": " + cudaGetErrorString(e_));                    \
            return -1;                                                                             \
        }                                                                                          \
    } while (0)

struct ScPrep {  // particle-independent part of a scatter call (scatter_prepare)
    hs::ScArgs a;
    int n_inc = 0, n_geom = 0, threads = 64;
    const int *inc_of_geom = nullptr;
    double *xs_inc = nullptr, *d_xs = nullptr;
};
static inline size_t sc_smem(int nto, int nmax_bound)
{
    // nto = orientations per pass (threads / SPLIT).  T stage + max(reduction rows, the per-orientation P' / Q' vectors
    // that alias them) + accumulators + per-orientation amplitude sums
    const size_t rows = (size_t)nto * (24 + 3), vecs = (size_t)2 * (nmax_bound + 2) * nto;
    return ((size_t)4 * nmax_bound * nmax_bound + std::max(rows, vecs) + SCK_G * 24 + 2 + (size_t)8 * SCK_G * nto) * sizeof(double);
}
// k_scatter<SPLIT> launch: `nto` orientations per pass x SPLIT lanes each
static inline cudaError_t sc_launch(int split, int grid, int nto, size_t smem, cudaStream_t st, const hs::ScArgs &a)
{
    if (nto == 64) {
        if (split == 4) hs::k_scatter<4, 64><<<grid, nto * 4, smem, st>>>(a);
        else if (split == 2) hs::k_scatter<2, 64><<<grid, nto * 2, smem, st>>>(a);
        else hs::k_scatter<1, 64><<<grid, nto, smem, st>>>(a);
    } else {
        if (split == 4) hs::k_scatter<4, 0><<<grid, nto * 4, smem, st>>>(a);
        else if (split == 2) hs::k_scatter<2, 0><<<grid, nto * 2, smem, st>>>(a);
        else hs::k_scatter<1, 0><<<grid, nto, smem, st>>>(a);
    }
    return cudaGetLastError();
}
static inline cudaError_t sc_set_smem(int bytes)
{
    const void *f[6] = {(const void *)hs::k_scatter<1, 64>, (const void *)hs::k_scatter<2, 64>, (const void *)hs::k_scatter<4, 64>,
                        (const void *)hs::k_scatter<1, 0>, (const void *)hs::k_scatter<2, 0>, (const void *)hs::k_scatter<4, 0>};
    for (int i = 0; i < 6; i++) {
        cudaError_t e = cudaFuncSetAttribute(f[i], cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}
static int ensure_gauss(hs_ctx *ctx, int gmax);
static int ensure_gauss_full(hs_ctx *ctx, int nmaxrule);
#include "hs_staged_host.h"
#include "hs_staged_dev.h"
struct StBufsHolder { StShared s; DvShared d; };

static int ensure_gauss(hs_ctx *ctx, int gmax)
{
    if (gmax <= ctx->gmax) return 0;
    if (gmax > HS_NPNG1) gmax = HS_NPNG1;
    size_t tot = (size_t)gmax * (gmax + 1) / 2;
    std::vector<double> hx(tot), hw(tot);
    for (int g = 1; g <= gmax; g++) hs::gauss_positive(2 * g, &hx[(size_t)g * (g - 1) / 2], &hw[(size_t)g * (g - 1) / 2]);
    double *nx = nullptr, *nw = nullptr;
    CK(cudaMalloc(&nx, tot * sizeof(double)));
    CK(cudaMalloc(&nw, tot * sizeof(double)));
    CK(cudaMemcpy(nx, hx.data(), tot * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(nw, hw.data(), tot * sizeof(double), cudaMemcpyHostToDevice));
    if (ctx->d_gx) ctx->gauss_old.push_back(ctx->d_gx);
    if (ctx->d_gw) ctx->gauss_old.push_back(ctx->d_gw);
    ctx->d_gx = nx; ctx->d_gw = nw; ctx->gmax = gmax;
    return 0;
}

// full n-point rules for the two-segment quadrature of cylinders (same retirement scheme as ensure_gauss: kernels in
// flight may still read the old tables)
static int ensure_gauss_full(hs_ctx *ctx, int nmaxrule)
{
    if (nmaxrule <= ctx->gfmax) return 0;
    nmaxrule = std::min(std::max(nmaxrule, 64), HS_NPNG1 / 2 + 1);
    size_t tot = (size_t)nmaxrule * (nmaxrule + 1) / 2;
    std::vector<double> hx(tot), hw(tot);
    for (int n = 1; n <= nmaxrule; n++) hs::gauss_full(n, &hx[(size_t)n * (n - 1) / 2], &hw[(size_t)n * (n - 1) / 2]);
    double *nx = nullptr, *nw = nullptr;
    CK(cudaMalloc(&nx, tot * sizeof(double)));
    CK(cudaMalloc(&nw, tot * sizeof(double)));
    CK(cudaMemcpy(nx, hx.data(), tot * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(nw, hw.data(), tot * sizeof(double), cudaMemcpyHostToDevice));
    if (ctx->d_gfx) ctx->gauss_old.push_back(ctx->d_gfx);
    if (ctx->d_gfw) ctx->gauss_old.push_back(ctx->d_gfw);
    ctx->d_gfx = nx; ctx->d_gfw = nw; ctx->gfmax = nmaxrule;
    return 0;
}

extern "C" const char *hs_version(void) { return "hydrotm-b200 0.1 (sm_100a)"; }
extern "C" long long hs_tmat_size(int nmax) { return hs::tmat_size(nmax); }
extern "C" long long hs_tmat_mode_offset(int nmax, int m)
{
    long long off = 0;
    for (int mm = 0; mm < m; mm++) { long long nm = nmax - (mm > 1 ? mm : 1) + 1; off += 2 * nm * nm; }
    return off;
}
extern "C" int hs_nmax_start(double r_ev, double wavelength)
{
    double xev = 2.0 * acos(-1.0) * r_ev / wavelength;
    int ixxx = (int)(xev + 4.05 * pow(xev, 0.333333));
    return ixxx > 4 ? ixxx : 4;
}

extern "C" int hs_ctx_create(hs_ctx **out, int device)
{
    hs_ctx *ctx = new hs_ctx();
    *out = ctx;
    ctx->device = device;
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = (int)prop.sharedMemPerBlockOptin;
    CK(cudaMalloc(&ctx->d_queue, 16 * sizeof(int)));
    for (int b = 0; b < 8; b++) {
        CK(cudaStreamCreateWithFlags(&ctx->bstream[b], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ctx->ev_done[b], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&ctx->ev_start, cudaEventDisableTiming));
    if (ensure_gauss(ctx, 64)) return -1;
    return 0;
}

extern "C" void hs_ctx_destroy(hs_ctx *ctx)
{
    if (!ctx) return;
    cudaFree(ctx->d_gx); cudaFree(ctx->d_gw); cudaFree(ctx->d_gfx); cudaFree(ctx->d_gfw); cudaFree(ctx->d_queue); cudaFree(ctx->d_items);
    cudaFree(ctx->d_ws_global); cudaFree(ctx->d_ws_fb); cudaFree(ctx->d_sc); cudaFree(ctx->d_xs_tmp); cudaFree(ctx->d_order); cudaFree(ctx->d_frames); cudaFree(ctx->d_scwig);
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    for (int b = 0; b < 8; b++) { if (ctx->bstream[b]) cudaStreamDestroy(ctx->bstream[b]); if (ctx->ev_done[b]) cudaEventDestroy(ctx->ev_done[b]); }
    if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
    if (ctx->st) { st_destroy(ctx->st->s); dv_destroy(ctx->st->d); delete ctx->st; }
    for (double *q : ctx->gauss_old) cudaFree(q);
    delete ctx;
}

extern "C" const char *hs_last_error(const hs_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" long long hs_launch_count(const hs_ctx *ctx) { return ctx ? ctx->launches : 0; }

static int ensure_pin(hs_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->h_pin_cap) return 0;
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    ctx->h_pin = nullptr; ctx->h_pin_cap = 0;
    CK(cudaMallocHost(&ctx->h_pin, bytes));
    ctx->h_pin_cap = bytes;
    return 0;
}

__global__ void k_xs_expand(int n, int n_geom, int n_inc, const int *inc_of_geom, const double *xs_inc, double *xs)
{
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)n * n_geom * 2) return;
    int c = (int)(i & 1);
    long pg = i >> 1;
    int g = (int)(pg % n_geom);
    long p = pg / n_geom;
    xs[i] = xs_inc[(p * n_inc + inc_of_geom[g]) * 2 + c];
}

// Everything of a scatter call that does not depend on the particles: geometry grouping, parameter block, the
// orientation / geometry frames.  Shared by hs_scatter_batch and the fused hs_calctmat_scatter_batch.
static int scatter_prepare(hs_ctx *ctx, int n, const double *d_tarena, const long long *d_toff, const int *d_nmax,
                           const int *d_status, const double *d_lam, int n_geom, const double *h_geom, int n_orient,
                           const double *d_alpha, const double *d_beta, const double *d_weight, double scale,
                           double *d_S, double *d_Z, double *d_xs, cudaStream_t st, ScPrep &P)
{
    if (n_geom > 64) { hs_set_err(ctx, "at most 64 geometries per call"); return -3; }
    // group geometries by incidence direction
    std::vector<double> inc;           // unique (thet0, phi0)
    std::vector<int> inc_of(n_geom);
    for (int g = 0; g < n_geom; g++) {
        int f = -1;
        for (size_t d = 0; d < inc.size() / 2; d++)
            if (inc[2 * d] == h_geom[4 * g] && inc[2 * d + 1] == h_geom[4 * g + 2]) { f = (int)d; break; }
        if (f < 0) { f = (int)(inc.size() / 2); inc.push_back(h_geom[4 * g]); inc.push_back(h_geom[4 * g + 2]); }
        inc_of[g] = f;
    }
    const int n_inc = (int)(inc.size() / 2);
    std::vector<int> g_begin(n_inc + 1, 0), g_out(n_geom);
    std::vector<double> gsc(2 * n_geom);
    int pos = 0;
    for (int d = 0; d < n_inc; d++) {
        g_begin[d] = pos;
        for (int g = 0; g < n_geom; g++)
            if (inc_of[g] == d) { gsc[2 * pos] = h_geom[4 * g + 1]; gsc[2 * pos + 1] = h_geom[4 * g + 3]; g_out[pos] = g; pos++; }
    }
    g_begin[n_inc] = pos;
    // pack: [inc 2*64][gsc 2*64][g_begin 65 ints][g_out 64 ints][inc_of 64 ints] in one small device block
    const size_t blk_doubles = 128 + 128 + 128;
    if (!ctx->d_sc) CK(cudaMalloc(&ctx->d_sc, blk_doubles * sizeof(double) * 4));
    static thread_local int slot = 0;
    slot = (slot + 1) & 3;  // 4 rotating parameter blocks so that back-to-back async calls do not overwrite each other
    std::vector<double> hostblk(blk_doubles, 0.0);
    memcpy(hostblk.data(), inc.data(), inc.size() * sizeof(double));
    memcpy(hostblk.data() + 128, gsc.data(), gsc.size() * sizeof(double));
    int *ib = reinterpret_cast<int *>(hostblk.data() + 256);
    memcpy(ib, g_begin.data(), g_begin.size() * sizeof(int));
    memcpy(ib + 65, g_out.data(), g_out.size() * sizeof(int));
    memcpy(ib + 130, inc_of.data(), inc_of.size() * sizeof(int));
    double *dblk = ctx->d_sc + (size_t)slot * blk_doubles;
    CK(cudaMemcpyAsync(dblk, hostblk.data(), blk_doubles * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));  // hostblk is pageable stack/heap memory: make the copy complete before it dies
    const int *dib = reinterpret_cast<const int *>(dblk + 256);
    double *xs_inc = nullptr;
    if (d_xs) {
        size_t need = (size_t)n * n_inc * 2 * sizeof(double);
        if (need > ctx->xs_tmp_cap) {
            if (ctx->d_xs_tmp) cudaFree(ctx->d_xs_tmp);
            ctx->d_xs_tmp = nullptr; ctx->xs_tmp_cap = 0;
            CK(cudaMalloc(&ctx->d_xs_tmp, need));
            ctx->xs_tmp_cap = need;
        }
        xs_inc = ctx->d_xs_tmp;
    }
    hs::ScArgs a;
    a.n = n; a.tarena = d_tarena; a.toff = d_toff; a.nmax = d_nmax; a.status = d_status; a.lam = d_lam;
    a.n_inc = n_inc; a.inc = dblk; a.g_begin = dib; a.gsc = dblk + 128; a.g_out = dib + 65; a.n_geom = n_geom;
    a.n_orient = n_orient; a.alpha = d_alpha; a.beta = d_beta; a.weight = d_weight; a.scale = scale;
    a.S = d_S; a.Z = d_Z; a.xs = xs_inc; a.wtab = nullptr; a.wig_nb = 0;
    {   // orientation / geometry frames: once per call instead of once per particle
        int maxblk = 1;
        for (int d = 0; d < n_inc; d++) maxblk = std::max(maxblk, (g_begin[d + 1] - g_begin[d] + SCK_G - 1) / SCK_G);
        size_t need = (size_t)n_inc * maxblk * n_orient * SC_FR;
        if (need > ctx->frames_cap) {
            if (ctx->d_frames) cudaFree(ctx->d_frames);
            ctx->d_frames = nullptr; ctx->frames_cap = 0;
            CK(cudaMalloc(&ctx->d_frames, need * sizeof(double)));
            ctx->frames_cap = need;
        }
        a.frames = ctx->d_frames; a.frame_blocks = maxblk;
        dim3 fg((n_orient + 63) / 64, n_inc, maxblk);
        hs::k_scatter_frames<<<fg, 64, 0, st>>>(a, ctx->d_frames);
        ctx->launches++;
    }
    P.a = a; P.n_inc = n_inc; P.inc_of_geom = dib + 130; P.xs_inc = xs_inc; P.d_xs = d_xs; P.n_geom = n_geom;
    P.threads = std::min(64, std::max(32, (n_orient + 31) / 32 * 32));
    return 0;
}

// Wigner table of a scatter call up to n = nb (the largest NMAX the launch will meet); a.frames must be in place
static int scatter_build_wig(hs_ctx *ctx, hs::ScArgs &a, int nb, cudaStream_t st)
{
    nb = std::max(1, std::min(nb, HS_NPN1));
    const int ndb = a.n_inc * a.frame_blocks;
    const size_t need = (size_t)ndb * (1 + SCK_G) * hs::sc_tri_size(nb) * a.n_orient;
    if (need > ctx->scwig_cap) {
        if (ctx->d_scwig) cudaFree(ctx->d_scwig);
        ctx->d_scwig = nullptr; ctx->scwig_cap = 0;
        CK(cudaMalloc(&ctx->d_scwig, need * sizeof(double2)));
        ctx->scwig_cap = need;
    }
    a.wtab = ctx->d_scwig; a.wig_nb = nb;
    dim3 g((a.n_orient + 63) / 64, nb + 1, ndb * (1 + SCK_G));
    hs::k_scatter_wig<<<g, 64, 0, st>>>(a, ctx->d_scwig);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

// bucket = launch configuration
struct Bucket {
    bool warp;        // group = warp (several particles per CTA) or whole CTA
    bool global_ws;
    int threads;
    int groups;       // groups per CTA
    int ws_cap;       // doubles per group
};

// Predicted converged NMAX per particle (bucket choice only; the kernel still starts from upstream's NMAX0).
// NMAX0 = max(4, int(x + 4.05 x^(1/3))) underestimates the converged order when |m| x is large (W band: 14 -> 25);
// max(NMAX0 + 1, ceil(|m| k a_max + 2)) is a tight estimate on the 6-band rain sweep; particles that still outgrow
// their shared-memory slice continue in the global-memory fall-back workspace (no re-launch).
__global__ void k_nmax0(int n, const double *axi, const double *lam, const double *mr, const double *mi,
                        const double *eps, double rat, int np, int *out, int *out_n0, int *out_lb)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double PI = 3.14159265358979323846;
    double r = rat;
    double e_ = eps[i];
    if (fabs(rat - 1.0) > 1e-8 && np == -2) {  // SAREAC
        r = pow(1.5 / e_, 1.0 / 3.0);
        r = r / sqrt((e_ + 2.0) / (2.0 * e_));
    } else if (fabs(rat - 1.0) > 1e-8) {
        double e, q;
        if (e_ < 1.0) { e = sqrt(1.0 - e_ * e_); q = 0.5 * (pow(e_, 2.0 / 3.0) + pow(e_, -1.0 / 3.0) * asin(e) / e); }
        else { e = sqrt(1.0 - 1.0 / (e_ * e_)); q = 0.25 * (2.0 * pow(e_, 2.0 / 3.0) + pow(e_, -4.0 / 3.0) * log((1.0 + e) / (1.0 - e)) / e); }
        r = 1.0 / sqrt(q);
    }
    double xev = 2.0 * PI * r * axi[i] / lam[i];
    int ixxx = (xev > 0.0 && xev < 1e6) ? (int)(xev + 4.05 * pow(xev, 0.333333)) : 0;
    int n0 = ixxx > 4 ? ixxx : 4;
    double a = (e_ > 0.0) ? pow(e_, 1.0 / 3.0) : 1.0;
    double xa = xev * fmax(a, a / e_);
    if (np == -2 && e_ > 0.0) xa = xev * pow(2.0 / (3.0 * e_ * e_), 1.0 / 3.0) * sqrt(1.0 + e_ * e_);  // edge of the cylinder
    double mx = sqrt(mr[i] * mr[i] + mi[i] * mi[i]) * xa;
    int pred = (mx > 0.0 && mx < 1e6) ? (int)ceil(mx + 2.0) : 0;
    out[i] = max(n0 + 1, pred);
    const bool bad = !(axi[i] > 0.0) || !(lam[i] > 0.0) || !(e_ > 0.0) || !isfinite(mr[i]) || !isfinite(mi[i]) || !(mr[i] * mr[i] + mi[i] * mi[i] > 0.0);
    out_n0[i] = bad ? -1 : n0;
    // lower-bound guess of the converged NMAX (speculation depth of the first round; only costs work if wrong):
    // NMAX ~ 0.60 |m| k a_max + 4.5 on the rain sweep (r = 0.98), minus a margin of 2
    out_lb[i] = (mx > 0.0 && mx < 1e6) ? max(n0 + 1, (int)floor(0.6035 * mx + 4.49) - 2) : n0 + 1;
}

static int print_staged_prof(hs_ctx *ctx)
{
    long long hp[24];
    CK(cudaMemcpy(hp, ctx->d_prof, sizeof(hp), cudaMemcpyDeviceToHost));
    {
            // staged pipeline: phases of k_st_assemble (thread 0 of every CTA)
            const char *nm[5] = {"prologue", "chunk loop", "epilogue", "solve", "T write"};
            double t5 = 0; for (int i = 0; i < 5; i++) t5 += (double)hp[i];
            const double ncta = (double)hp[5] > 0 ? (double)hp[5] : 1.0;
            for (int i = 0; i < 5; i++)
                fprintf(stderr, "[hydrotm-prof] k_st_assemble %-10s %6.2f%%  %8.0f cycles / CTA\n", nm[i], 100.0 * hp[i] / t5, (double)hp[i] / ncta);
            fprintf(stderr, "[hydrotm-prof] k_st_assemble chunk loop split: build %.0f, barrier %.0f, contraction %.0f, barrier %.0f cycles / CTA\n",
                    (double)hp[8] / ncta, (double)hp[9] / ncta, (double)hp[10] / ncta, (double)hp[11] / ncta);
            fprintf(stderr, "[hydrotm-prof] k_st_assemble solve: %.1f elimination steps / CTA -> %.0f cycles / step\n", (double)hp[12] / ncta, (double)hp[3] / std::max(1.0, (double)hp[12]));
            {
                const double nw = (double)hp[21] > 0 ? (double)hp[21] : 1.0;
                fprintf(stderr, "[hydrotm-prof] k_st_mode_small %.0f warps (mean NM %.1f, g %.1f): prologue %.0f, node loop %.0f, combine %.0f, solve %.0f, T write %.0f cycles / warp\n",
                        nw, (double)hp[22] / nw, (double)hp[23] / nw, (double)hp[16] / nw, (double)hp[17] / nw, (double)hp[18] / nw, (double)hp[19] / nw, (double)hp[20] / nw);
            }
            fprintf(stderr, "[hydrotm-prof] k_st_assemble %.0f CTAs, mean NM %.1f, mean g %.1f, %.0f cycles / CTA\n", ncta, (double)hp[6] / ncta, (double)hp[7] / ncta, t5 / ncta);
    }
    return 0;
}

template <bool WARP, bool GLOBAL_WS, bool TILED>
static cudaError_t launch_tm(const hs::TmArgs &a, int grid, int threads, size_t smem, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(hs::k_calctmat<WARP, GLOBAL_WS, TILED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    hs::k_calctmat<WARP, GLOBAL_WS, TILED><<<grid, threads, smem, st>>>(a);
    return cudaGetLastError();
}

// sp != null: fused scatter plan; *sp_used tells the caller whether the staged pipeline consumed it (every particle staged)
static int calctmat_impl(hs_ctx *ctx, int n, const double *d_axi, const double *d_lam, const double *d_mr,
                         const double *d_mi, const double *d_eps, double rat, int np, double ddelt,
                         int ndgs, int rescale_ddelt, double *d_tarena, long long arena_cap,
                         unsigned long long *d_arena_top, long long *d_toff, int *d_nmax, int *d_ngauss,
                         int *d_ntrials, int *d_status, void *stream, const ScPrep *sp, bool *sp_used)
{
    if (!ctx) return -2;
    if (n <= 0) return 0;
    if (np != HS_SHAPE_SPHEROID && np != HS_SHAPE_CYLINDER) { hs_set_err(ctx, "shapes: spheroid (np = -1) or finite circular cylinder (np = -2)"); return -3; }
    if (np == HS_SHAPE_CYLINDER && getenv("HYDROTM_STAGED_MIN") && atoi(getenv("HYDROTM_STAGED_MIN")) > 0) {
        hs_set_err(ctx, "cylinders are only implemented in the staged pipeline (unset HYDROTM_STAGED_MIN)"); return -3;
    }
    if (ndgs < 1) { hs_set_err(ctx, "ndgs must be >= 1"); return -3; }
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    if (rescale_ddelt) ddelt = 0.1 * ddelt;
    const bool trace = getenv("HYDROTM_TRACE") != nullptr;
    auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!trace) return;
        auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[hydrotm] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - T0).count());
    };

    // Size buckets (launch configurations), smallest first; a particle that outgrows its bucket is re-run in the
    // next one.  Bucket i packs k_i CTAs per SM: its per-particle workspace is the k_i-th share of the 227 KB of
    // shared memory.  Small CTAs + many particles in flight beat big CTAs here: the per-particle critical path
    // (recurrences, pivot searches, barriers) does not shrink with more threads.
    // Override for experiments: HYDROTM_BUCKETS="k:threads[:w],..." (w = warp-per-particle groups).
    Bucket buckets[8];
    int NB = 0;
    {
        int ks[7] = {8, 5, 3, 2, 1, 0, 0}, th[7] = {64, 128, 128, 128, 256, 0, 0}, wg[7] = {0, 0, 0, 0, 0, 0, 0};
        int nk = 5;
        const char *env = getenv("HYDROTM_BUCKETS");
        if (env && *env) {
            nk = 0;
            const char *q = env;
            while (*q && nk < 7) {
                int k = 0, t = 0, w = 0, used = 0;
                if (sscanf(q, "%d:%d%n", &k, &t, &used) < 2) break;
                q += used;
                if (*q == ':') { w = 1; q += 2; }
                ks[nk] = k; th[nk] = t; wg[nk] = w; nk++;
                if (*q == ',') q++;
            }
        }
        for (int i = 0; i < nk; i++) {
            int k = std::max(1, ks[i]);
            // 228 KB per SM; each resident CTA also needs ~7 KB static smem (control blocks, tables) + 1 KB reserved
            int per_cta = std::min(ctx->smem_optin - 8192, (233472 - k * 9216) / k);  // bytes of dynamic smem per CTA
            Bucket B;
            B.global_ws = false;
            B.warp = wg[i] != 0;
            B.threads = th[i];
            B.groups = B.warp ? std::max(1, th[i] / 32) : 1;
            if (k == 1 && getenv("HYDROTM_HEAVY_CAP_KB")) per_cta = std::min(per_cta, atoi(getenv("HYDROTM_HEAVY_CAP_KB")) * 1024);
            B.ws_cap = (per_cta / 8 / B.groups) & ~1;
            buckets[NB++] = B;
        }
        Bucket G = {false, true, 256, 1, (int)((hs::ws_need(HS_NPN1, HS_NPNG1) + 1) & ~1LL)};
        buckets[NB++] = G;
    }

    // fall-back workspace: one slice per resident group (<= 8 groups per SM), sized for NMAX 48 at NDGS 2
    const int fb_cap = (int)(hs::ws_need(48, 97) + 1) & ~1;
    if (!ctx->d_ws_fb) {
        if (cudaMalloc(&ctx->d_ws_fb, (size_t)ctx->sm_count * 8 * fb_cap * sizeof(double)) != cudaSuccess) { ctx->d_ws_fb = nullptr; cudaGetLastError(); }
    }
    // Device-driven rounds (hs_staged_dev.h): controller, planner and task lists on the device, no host round trip between
    // rounds.  Default for batches of HYDROTM_DEVCTL_MIN (256) particles or more; HYDROTM_DEVCTL=0 keeps the host-driven
    // rounds (hs_staged_host.h), which small calls use anyway (a handful of particles is latency-bound either way).
    {
        const int devctl = getenv("HYDROTM_DEVCTL") ? atoi(getenv("HYDROTM_DEVCTL")) : 1;
        const int devmin = getenv("HYDROTM_DEVCTL_MIN") ? atoi(getenv("HYDROTM_DEVCTL_MIN")) : 256;
        const bool fused_req = getenv("HYDROTM_STAGED_MIN") && atoi(getenv("HYDROTM_STAGED_MIN")) > 0;
        const bool host_only = (getenv("HYDROTM_FUSE_SOLVE") && atoi(getenv("HYDROTM_FUSE_SOLVE")) == 0) ||
                               (getenv("HYDROTM_KR16") && atoi(getenv("HYDROTM_KR16")) == 0) || getenv("HYDROTM_OVERLAP_SCATTER") ||
                               getenv("HYDROTM_TIMELINE");
        if (devctl && n >= devmin && !fused_req && !host_only) {
            if (!ctx->st) ctx->st = new StBufsHolder();
            if (getenv("HYDROTM_PROF") && !ctx->d_prof) { CK(cudaMalloc(&ctx->d_prof, 24 * sizeof(long long))); }
            if (ctx->d_prof) CK(cudaMemsetAsync(ctx->d_prof, 0, 24 * sizeof(long long), st));
            StCall SC;
            SC.d_axi = d_axi; SC.d_lam = d_lam; SC.d_mr = d_mr; SC.d_mi = d_mi; SC.d_eps = d_eps; SC.rat = rat; SC.np = np; SC.ddelt = ddelt; SC.ndgs = ndgs;
            SC.d_tarena = d_tarena; SC.arena_cap = arena_cap; SC.d_arena_top = d_arena_top; SC.d_toff = d_toff;
            SC.d_nmax = d_nmax; SC.d_ngauss = d_ngauss; SC.d_ntrials = d_ntrials; SC.d_status = d_status;
            SC.scratch_budget = (size_t)((getenv("HYDROTM_SCRATCH_GB") ? atof(getenv("HYDROTM_SCRATCH_GB")) : 4.0) * (double)(1ull << 30));
            SC.spec = getenv("HYDROTM_SPEC") ? atoi(getenv("HYDROTM_SPEC")) : 0;
            SC.trace = trace;
            SC.prof = (unsigned long long *)ctx->d_prof;
            SC.one_lane = getenv("HYDROTM_ONE_LANE") && atoi(getenv("HYDROTM_ONE_LANE")) != 0;
            const int rc = calctmat_dev(ctx, ctx->st->s, ctx->st->d, st, n, SC);
            if (rc) return rc;
            lap("device-driven rounds");
            if (ctx->d_prof) print_staged_prof(ctx);
            if (SC.nmax_max > ctx->last_nmax_max) ctx->last_nmax_max = SC.nmax_max;
            return SC.arena_full ? HS_RC_ARENA_FULL : 0;
        }
    }
    if (ensure_pin(ctx, (size_t)n * sizeof(int) * 4)) return -1;
    if (n > ctx->items_cap) {
        if (ctx->d_items) cudaFree(ctx->d_items);
        ctx->d_items = nullptr; ctx->items_cap = 0;
        CK(cudaMalloc(&ctx->d_items, (size_t)n * sizeof(int)));
        ctx->items_cap = n;
    }
    // predicted starting NMAX -> initial bucket
    // (the kernel writes straight into the pinned host block: no D2H copy that could queue behind a caller's bulk transfer)
    int *h_n0 = ctx->h_pin, *h_st = ctx->h_pin + n, *h_start = ctx->h_pin + 2 * (size_t)n, *h_lb = ctx->h_pin + 3 * (size_t)n;
    k_nmax0<<<(n + 255) / 256, 256, 0, st>>>(n, d_axi, d_lam, d_mr, d_mi, d_eps, rat, np, h_n0, h_start, h_lb);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    // Particles whose predicted order reaches HYDROTM_STAGED_MIN go through the staged pipeline (one kernel per stage,
    // trials and modes as independent work items); the small ones stay on the fused one-particle-per-group kernel.
    const int staged_min = getenv("HYDROTM_STAGED_MIN") ? atoi(getenv("HYDROTM_STAGED_MIN")) : 0;
    std::vector<int> st_sel, st_n0, st_lb;
    st_sel.reserve(n); st_n0.reserve(n); st_lb.reserve(n);
    if (!ctx->st) ctx->st = new StBufsHolder();
    bool staged_done = false;
    StCall SC;
    SC.d_axi = d_axi; SC.d_lam = d_lam; SC.d_mr = d_mr; SC.d_mi = d_mi; SC.d_eps = d_eps; SC.rat = rat; SC.np = np; SC.ddelt = ddelt; SC.ndgs = ndgs;
    SC.d_tarena = d_tarena; SC.arena_cap = arena_cap; SC.d_arena_top = d_arena_top; SC.d_toff = d_toff;
    SC.d_nmax = d_nmax; SC.d_ngauss = d_ngauss; SC.d_ntrials = d_ntrials; SC.d_status = d_status;
    SC.scratch_budget = (size_t)((getenv("HYDROTM_SCRATCH_GB") ? atof(getenv("HYDROTM_SCRATCH_GB")) : 4.0) * (double)(1ull << 30));
    SC.spec = getenv("HYDROTM_SPEC") ? atoi(getenv("HYDROTM_SPEC")) : 0;
    SC.trace = trace;
    SC.sp = nullptr;
    SC.fuse_solve = !(getenv("HYDROTM_FUSE_SOLVE") && atoi(getenv("HYDROTM_FUSE_SOLVE")) == 0);
    SC.prof = (unsigned long long *)ctx->d_prof;
    SC.kr16 = !(getenv("HYDROTM_KR16") && atoi(getenv("HYDROTM_KR16")) == 0);
    SC.one_lane = getenv("HYDROTM_ONE_LANE") && atoi(getenv("HYDROTM_ONE_LANE")) != 0;
    auto run_staged = [&]() -> int {
        staged_done = true;
        return calctmat_staged(ctx, ctx->st->s, st, st_sel, st_n0, st_lb, SC);
    };

    if (getenv("HYDROTM_PROF") && !ctx->d_prof) { CK(cudaMalloc(&ctx->d_prof, 24 * sizeof(long long))); }
    if (ctx->d_prof) CK(cudaMemsetAsync(ctx->d_prof, 0, 24 * sizeof(long long), st));
    lap("nmax0 kernel + D2H");
    std::vector<int> bucket_of(n), nm0(n);
    std::vector<std::vector<int>> lists(NB);
    for (int i = 0; i < n; i++) {
        int n0 = h_n0[i];
        nm0[i] = n0;
        if (n0 >= staged_min && h_start[i] >= 0) { st_sel.push_back(i); st_n0.push_back(h_start[i]); st_lb.push_back(h_lb[i]); bucket_of[i] = -1; continue; }
        int np_ = std::min(n0, HS_NPN1);
        long long need = hs::ws_need(np_, np_ * ndgs + 1);
        int b = 0;
        while (b < NB - 1 && need > buckets[b].ws_cap) b++;
        // the last shared-memory bucket also takes particles predicted slightly beyond its slice: they start in smem
        // (the trial sequence grows from NMAX0) and only move to the fall-back workspace if they really outgrow it
        if (b == NB - 1 && NB >= 2 && ctx->d_ws_fb && need <= fb_cap) b = NB - 2;
        bucket_of[i] = b;
        lists[b].push_back(i);
    }
    // Overlapping the scatter kernels of finished particles with the convergence rounds of the others measured SLOWER on
    // B200 (12.3 vs 11.6 ms per W2 step: both sides are limited by resident-CTA resources, not by idle SMs), so the
    // overlap is opt-in (HYDROTM_OVERLAP_SCATTER=1) and the default runs the two stages back to back.
    if (sp && (int)st_sel.size() == n && getenv("HYDROTM_OVERLAP_SCATTER")) { SC.sp = sp; if (sp_used) *sp_used = true; }
    int gauss_needed = ctx->gmax;
    for (int round = 0; round < 16; round++) {
        bool any = false;
        std::vector<int> order;
        order.reserve(n);
        std::vector<int> seg_start(NB + 1, 0);
        for (int b = 0; b < NB; b++) {
            auto &L = lists[b];
            std::stable_sort(L.begin(), L.end(), [&](int x, int y) { return nm0[x] > nm0[y]; });
            seg_start[b] = (int)order.size();
            order.insert(order.end(), L.begin(), L.end());
            if (!L.empty()) any = true;
        }
        seg_start[NB] = (int)order.size();
        if (!any) break;
        lap("host bucketing/sort");
        CK(cudaMemcpyAsync(ctx->d_items, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(ctx->d_queue, 0, 16 * sizeof(int), st));
        CK(cudaEventRecord(ctx->ev_start, st));
        for (int b = NB - 1; b >= 0; b--) {  // heaviest bucket first; buckets run concurrently on their own streams
            int cnt = seg_start[b + 1] - seg_start[b];
            if (!cnt) continue;
            const Bucket &B = buckets[b];
            hs::TmArgs a;
            a.items = ctx->d_items + seg_start[b]; a.n_items = cnt; a.queue = ctx->d_queue + b;
            a.axi = d_axi; a.lam = d_lam; a.mr = d_mr; a.mi = d_mi; a.eps = d_eps;
            a.rat = rat; a.np = np; a.ddelt = ddelt; a.ndgs = ndgs;
            a.gx = ctx->d_gx; a.gw = ctx->d_gw; a.gmax = ctx->gmax;
            a.tarena = d_tarena; a.arena_cap = arena_cap; a.arena_top = d_arena_top;
            a.toff = d_toff; a.nmax = d_nmax; a.ngauss = d_ngauss; a.ntrials = d_ntrials; a.status = d_status;
            a.ws_cap = B.ws_cap; a.ws_global = nullptr;
            a.prof = ctx->d_prof;
            a.tile_min = getenv("HYDROTM_TILE_MIN") ? atoi(getenv("HYDROTM_TILE_MIN")) : 12;
            a.debug = getenv("HYDROTM_DEBUG") ? atoi(getenv("HYDROTM_DEBUG")) : 0;
            a.ws_fb = B.global_ws ? nullptr : ctx->d_ws_fb; a.fb_cap = B.global_ws ? 0 : fb_cap;
            size_t smem = B.global_ws ? 0 : (size_t)B.groups * B.ws_cap * sizeof(double);
            int ctas_needed = (cnt + B.groups - 1) / B.groups;
            int per_sm = B.global_ws ? 1 : std::max(1, (int)((size_t)(ctx->smem_optin + 1024) / (smem + 4096)));
            int grid = std::min(ctas_needed, ctx->sm_count * std::min(per_sm, 8 / B.groups));
            cudaError_t e;
            cudaStream_t bs = ctx->bstream[b];
            CK(cudaStreamWaitEvent(bs, ctx->ev_start, 0));
            if (B.global_ws) {
                grid = std::min(ctas_needed, ctx->sm_count);
                size_t bytes = (size_t)grid * B.ws_cap * sizeof(double);
                if (bytes > ctx->ws_global_bytes) {
                    if (ctx->d_ws_global) cudaFree(ctx->d_ws_global);
                    ctx->d_ws_global = nullptr; ctx->ws_global_bytes = 0;
                    CK(cudaMalloc(&ctx->d_ws_global, bytes));
                    ctx->ws_global_bytes = bytes;
                }
                a.ws_global = ctx->d_ws_global;
                e = launch_tm<false, true, false>(a, grid, B.threads, 0, bs);
            } else if (B.warp) {
                e = launch_tm<true, false, false>(a, grid, B.threads, smem, bs);
            } else {
                e = (per_sm <= 2) ? launch_tm<false, false, true>(a, grid, B.threads, smem, bs)
                                 : launch_tm<false, false, false>(a, grid, B.threads, smem, bs);
            }
            ctx->launches++;
            if (e != cudaSuccess) { hs_set_err(ctx, std::string("k_calctmat launch: ") + cudaGetErrorString(e)); return -1; }
            CK(cudaEventRecord(ctx->ev_done[b], bs));
            CK(cudaStreamWaitEvent(st, ctx->ev_done[b], 0));
        }
        lap("launches issued");
        if (!staged_done) { int rc = run_staged(); if (rc) return rc; lap("staged pipeline"); }
        CK(cudaMemcpyAsync(h_st, d_status, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        lap("round done (status D2H)");
        // escalations
        std::vector<std::vector<int>> next(NB);
        bool need_gauss = false;
        for (int b = 0; b < NB; b++)
            for (int i : lists[b]) {
                if (h_st[i] == HS_ST_INTERNAL_ESCALATE) {
                    int nb = std::min(b + 1, NB - 1);
                    if (nb == b) { hs_set_err(ctx, "workspace escalation beyond the largest bucket"); return -4; }
                    next[nb].push_back(i);
                } else if (h_st[i] == HS_ST_INTERNAL_GAUSS) {
                    need_gauss = true;
                    next[b].push_back(i);
                }
            }
        if (need_gauss) {
            gauss_needed = std::min(HS_NPNG1, std::max(2 * ctx->gmax, 128));
            if (ensure_gauss(ctx, gauss_needed)) return -1;
        }
        lists.swap(next);
    }
    if (!staged_done) { int rc = run_staged(); if (rc) return rc; lap("staged pipeline (alone)"); }
    if (ctx->d_prof) {
        long long hp[24];
        CK(cudaMemcpy(hp, ctx->d_prof, sizeof(hp), cudaMemcpyDeviceToHost));
        double tot = 0; for (int i = 0; i < 8; i++) tot += (double)hp[i];
        if (st_sel.empty()) {
            const char *nm[8] = {"other", "trial_setup", "assemble m=0", "solve m=0", "mode_setup", "assemble m>=1", "solve m>=1", "write T"};
            for (int i = 0; i < 8; i++) fprintf(stderr, "[hydrotm-prof] %-14s %6.2f%%  %.3e cycles\n", nm[i], 100.0 * hp[i] / tot, (double)hp[i]);
        } else {
            for (int i = 0; i < 24; i++) (void)hp[i];
            print_staged_prof(ctx);
        }
    }
    // remember the largest converged NMAX (bounds the T stage of the scatter kernel); the staged pipeline reports its own
    if (SC.nmax_max > ctx->last_nmax_max) ctx->last_nmax_max = SC.nmax_max;
    bool arena_full = SC.arena_full;
    if ((int)st_sel.size() < n) {
        CK(cudaMemcpyAsync(h_n0, d_nmax, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h_st, d_status, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (int i = 0; i < n; i++) {
            if (h_n0[i] > ctx->last_nmax_max && h_n0[i] <= HS_NPN1) ctx->last_nmax_max = h_n0[i];
            if (h_st[i] == HS_ST_ARENA_FULL) arena_full = true;
        }
    }
    return arena_full ? HS_RC_ARENA_FULL : 0;
}

extern "C" int hs_calctmat_batch(hs_ctx *ctx, int n, const double *d_axi, const double *d_lam, const double *d_mr,
                                 const double *d_mi, const double *d_eps, double rat, int np, double ddelt,
                                 int ndgs, int rescale_ddelt, double *d_tarena, long long arena_cap,
                                 unsigned long long *d_arena_top, long long *d_toff, int *d_nmax, int *d_ngauss,
                                 int *d_ntrials, int *d_status, void *stream)
{
    return calctmat_impl(ctx, n, d_axi, d_lam, d_mr, d_mi, d_eps, rat, np, ddelt, ndgs, rescale_ddelt, d_tarena, arena_cap, d_arena_top,
                         d_toff, d_nmax, d_ngauss, d_ntrials, d_status, stream, nullptr, nullptr);
}

extern "C" int hs_scatter_batch(hs_ctx *ctx, int n, const double *d_tarena, const long long *d_toff, const int *d_nmax,
                                const int *d_status, const double *d_lam, int n_geom, const double *h_geom, int n_orient,
                                const double *d_alpha, const double *d_beta, const double *d_weight, double scale,
                                double *d_S, double *d_Z, double *d_xs, void *stream);

extern "C" int hs_calctmat_scatter_batch(hs_ctx *ctx, int n, const double *d_axi, const double *d_lam, const double *d_mr,
                                         const double *d_mi, const double *d_eps, double rat, int np, double ddelt, int ndgs,
                                         int rescale_ddelt, double *d_tarena, long long arena_cap, unsigned long long *d_arena_top,
                                         long long *d_toff, int *d_nmax, int *d_ngauss, int *d_ntrials, int *d_status, int n_geom,
                                         const double *h_geom, int n_orient, const double *d_alpha, const double *d_beta,
                                         const double *d_weight, double scale, double *d_S, double *d_Z, double *d_xs, void *stream)
{
    if (!ctx) return -2;
    if (n <= 0) return 0;
    if (n_geom <= 0 || n_orient <= 0) { hs_set_err(ctx, "hs_calctmat_scatter_batch: no geometries / orientations"); return -3; }
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    ScPrep SP;
    { int rc = scatter_prepare(ctx, n, d_tarena, d_toff, d_nmax, d_status, d_lam, n_geom, h_geom, n_orient, d_alpha, d_beta,
                               d_weight, scale, d_S, d_Z, d_xs, st, SP); if (rc) return rc; }
    CK(sc_set_smem(ctx->smem_optin - 8192));
    // scatter launches that overlap the convergence rounds do not know the largest NMAX yet: table up to the limit
    if (getenv("HYDROTM_OVERLAP_SCATTER")) { int rc = scatter_build_wig(ctx, SP.a, HS_NPN1, st); if (rc) return rc; }
    bool used = false;
    int rc = calctmat_impl(ctx, n, d_axi, d_lam, d_mr, d_mi, d_eps, rat, np, ddelt, ndgs, rescale_ddelt, d_tarena, arena_cap, d_arena_top,
                           d_toff, d_nmax, d_ngauss, d_ntrials, d_status, stream, &SP, &used);
    if (rc != 0) return rc;  // incl. HS_RC_ARENA_FULL: the caller retries with a larger arena
    if (!used)  // some particles went through the fused calctmat kernel: scatter afterwards
        return hs_scatter_batch(ctx, n, d_tarena, d_toff, d_nmax, d_status, d_lam, n_geom, h_geom, n_orient, d_alpha, d_beta, d_weight, scale,
                                d_S, d_Z, d_xs, stream);
    if (d_xs) {
        long tot = (long)n * n_geom * 2;
        k_xs_expand<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(n, n_geom, SP.n_inc, SP.inc_of_geom, SP.xs_inc, d_xs);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    return 0;
}

extern "C" int hs_calcampl_batch(hs_ctx *ctx, int n, const double *d_tarena, const long long *d_toff,
                                 const int *d_nmax, const int *d_status, const double *d_lam, int n_geom,
                                 const double *d_geom, int n_orient, const double *d_alpha, const double *d_beta,
                                 const double *d_weight, double scale, int reduce, double *d_S, double *d_Z,
                                 void *stream)
{
    if (!ctx) return -2;
    if (n <= 0 || n_geom <= 0 || n_orient <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    hs::AmplArgs a;
    a.n = n; a.tarena = d_tarena; a.toff = d_toff; a.nmax = d_nmax; a.status = d_status; a.lam = d_lam;
    a.n_geom = n_geom; a.geom = d_geom; a.n_orient = n_orient; a.alpha = d_alpha; a.beta = d_beta;
    a.weight = d_weight; a.scale = scale; a.reduce = reduce; a.S = d_S; a.Z = d_Z;
    int items = reduce ? n_orient : n_geom * n_orient;
    int threads = std::min(128, std::max(32, (items + 31) / 32 * 32));
    size_t smem = (size_t)threads * 25 * sizeof(double);
    hs::k_ampl<<<n, threads, smem, st>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int hs_xsect_batch(hs_ctx *ctx, int n, const double *d_tarena, const long long *d_toff, const int *d_nmax,
                              const int *d_status, const double *d_lam, int n_inc, const double *d_inc, int n_orient,
                              const double *d_alpha, const double *d_beta, const double *d_weight, double scale,
                              double *d_out, void *stream)
{
    if (!ctx) return -2;
    if (n <= 0 || n_inc <= 0 || n_orient <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    hs::XsArgs a;
    a.n = n; a.tarena = d_tarena; a.toff = d_toff; a.nmax = d_nmax; a.status = d_status; a.lam = d_lam;
    a.n_inc = n_inc; a.inc = d_inc; a.n_orient = n_orient; a.alpha = d_alpha; a.beta = d_beta; a.weight = d_weight;
    a.scale = scale; a.out = d_out;
    int threads = std::min(128, std::max(32, (n_orient + 31) / 32 * 32));
    hs::k_xsect<<<n, threads, (size_t)threads * 2 * sizeof(double), st>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int hs_psd_weights(hs_ctx *ctx, int kind, int n_dsd, const double *d_p0, const double *d_p1,
                              const double *d_p2, const double *d_dmax, int n_D, const double *d_D,
                              const double *d_wq, double *d_W, void *stream)
{
    if (!ctx) return -2;
    if (n_dsd <= 0 || n_D <= 0) return 0;
    if (kind < 0 || kind > 2) { hs_set_err(ctx, "unknown PSD kind"); return -3; }
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    long tot = (long)n_dsd * n_D;
    hs::k_psd_weights<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(kind, n_dsd, d_p0, d_p1, d_p2, d_dmax, n_D, d_D, d_wq, d_W);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int hs_psd_integrate(hs_ctx *ctx, int n_dsd, int n_D, int ncol, const double *d_W, const double *d_Tt,
                                double *d_out, void *stream)
{
    if (!ctx) return -2;
    if (n_dsd <= 0 || n_D <= 0 || ncol <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    dim3 grid((ncol + PSD_CB - 1) / PSD_CB, (n_dsd + PSD_RB - 1) / PSD_RB);
    hs::k_psd_gemm<<<grid, 256, 0, st>>>(n_dsd, n_D, ncol, d_W, d_Tt, d_out);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int hs_observables(hs_ctx *ctx, int n, int stride, const double *d_in, int off_back, int off_forw,
                              int off_ext, int off_sca, const double *d_lam, int lam_stride, double kw_sqr,
                              double *d_out, void *stream)
{
    if (!ctx) return -2;
    if (n <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    hs::k_observables<<<(n + 127) / 128, 128, 0, st>>>(n, stride, d_in, off_back, off_forw, off_ext, off_sca, d_lam,
                                                       lam_stride, kw_sqr, d_out);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}


static int psd_obs_impl(hs_ctx *ctx, int n_dsd, int n_D, int n_tab, const double *d_W, const double *d_rows, const double *d_lam_tab,
                        double kw_sqr, int hpol_quirk, int force_sca, int n_sel, const int *h_sel, double *d_obs, void *stream)
{
    if (!ctx) return -2;
    if (n_dsd <= 0 || n_D <= 0 || n_tab <= 0) return 0;
    if (n_sel < 1 || n_sel > 15 || !h_sel) { hs_set_err(ctx, "hs_psd_observables_sel: 1..15 selected columns"); return -3; }
    hs::PsdSel sel;
    sel.n = n_sel;
    int want_sca = force_sca > 0;
    for (int q = 0; q < 15; q++) sel.col[q] = 0;
    for (int q = 0; q < n_sel; q++) {
        if (h_sel[q] < 0 || h_sel[q] > 14) { hs_set_err(ctx, "hs_psd_observables_sel: column index outside 0..14"); return -3; }
        sel.col[q] = (signed char)h_sel[q];
        if (h_sel[q] < 2 && force_sca < 0) want_sca = 1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    // weights + one area that is in turn the table tile, the integrated values ([64][65]) and the staged observables
    // (the DMMA variant stages the table tile transposed: [64 columns][n_D | 1])
    static const bool dmma = !(getenv("HYDROTM_PSD_DMMA") && atoi(getenv("HYDROTM_PSD_DMMA")) == 0);
    const size_t tile = dmma ? (size_t)PSDF_TB * PSDF_NC * (n_D | 1) : (size_t)n_D * PSDF_TB * PSDF_NC;
    const size_t smem = ((size_t)PSDF_RB * (n_D | 1) + std::max(tile, (size_t)PSDF_RB * 65)) * sizeof(double);
    if (smem > (size_t)ctx->smem_optin) { hs_set_err(ctx, "hs_psd_observables: too many diameters for one shared-memory tile"); return -3; }
    CK(cudaFuncSetAttribute(hs::k_psd_obs_fused<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(hs::k_psd_obs_fused<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int rb = (n_dsd + PSDF_RB - 1) / PSDF_RB;
    // enough CTAs for two waves of the SMs: split the tables when there are few DSD tiles
    int splits = std::max(1, std::min((n_tab + PSDF_TB - 1) / PSDF_TB, (2 * ctx->sm_count + rb - 1) / rb));
    int per = ((n_tab + splits - 1) / splits + PSDF_TB - 1) / PSDF_TB * PSDF_TB;
    splits = (n_tab + per - 1) / per;
    dim3 grid(rb, splits);
    if (dmma) hs::k_psd_obs_fused<true><<<grid, 256, smem, st>>>(n_dsd, n_D, n_tab, per, d_W, d_rows, d_lam_tab, kw_sqr, hpol_quirk, want_sca, sel, d_obs);
    else hs::k_psd_obs_fused<false><<<grid, 256, smem, st>>>(n_dsd, n_D, n_tab, per, d_W, d_rows, d_lam_tab, kw_sqr, hpol_quirk, want_sca, sel, d_obs);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int hs_psd_observables(hs_ctx *ctx, int n_dsd, int n_D, int n_tab, const double *d_W, const double *d_rows,
                                  const double *d_lam_tab, double kw_sqr, int hpol_quirk, int want_sca, double *d_obs, void *stream)
{
    // all 15 columns; without want_sca the two scattering cross sections are NaN (the table build may not have them)
    int sel[15];
    for (int q = 0; q < 15; q++) sel[q] = q;
    return psd_obs_impl(ctx, n_dsd, n_D, n_tab, d_W, d_rows, d_lam_tab, kw_sqr, hpol_quirk, want_sca ? 1 : 0, 15, sel, d_obs, stream);
}

extern "C" int hs_psd_observables_sel(hs_ctx *ctx, int n_dsd, int n_D, int n_tab, const double *d_W, const double *d_rows,
                                      const double *d_lam_tab, double kw_sqr, int hpol_quirk, int n_sel, const int *h_sel,
                                      double *d_obs, void *stream)
{
    return psd_obs_impl(ctx, n_dsd, n_D, n_tab, d_W, d_rows, d_lam_tab, kw_sqr, hpol_quirk, -1, n_sel, h_sel, d_obs, stream);
}

extern "C" int hs_pack_rows(hs_ctx *ctx, int n, int n_geom, const double *d_S, const double *d_Z, const double *d_xs, const double *d_lam,
                            int g_back, int g_forw, int f_back, int f_forw, double *d_rows, void *stream)
{
    if (!ctx) return -2;
    if (n <= 0) return 0;
    const int idx[4] = {g_back, g_forw, f_back, f_forw};
    for (int q = 0; q < 4; q++)
        if (idx[q] < 0 || idx[q] >= n_geom) { hs_set_err(ctx, "hs_pack_rows: geometry index out of range"); return -3; }
    CK(cudaSetDevice(ctx->device));
    const long tot = (long)n * 56;
    hs::k_pack_rows<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, n_geom, d_S, d_Z, d_xs, d_lam, g_back, g_forw,
                                                                                    f_back, f_forw, d_rows);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int hs_gather_rows(hs_ctx *ctx, long long n_out, int ncol, const double *d_src, const long long *d_index, double *d_dst,
                              void *stream)
{
    if (!ctx) return -2;
    if (n_out <= 0 || ncol <= 0) return 0;
    CK(cudaSetDevice(ctx->device));
    const long tot = (long)n_out * ncol;
    hs::k_gather_rows<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>((long)n_out, ncol, d_src, d_index, d_dst);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

// ---- Monte-Carlo canting moments on numpy's random stream (hs_cantmc.cuh) ----
static long long cmc_nblocks(int n_cant, int n_part, double slack)
{
    return (long long)std::ceil((double)n_cant * (double)n_part * slack / CMC_B) + 8;
}
static size_t cmc_align(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" long long hs_canting_mc_scratch_bytes(int n_cant, int n_part, double slack)
{
    if (n_cant <= 0 || n_part <= 0) return 0;
    const long long nb = cmc_nblocks(n_cant, n_part, slack);
    size_t tot = 2 * cmc_align((size_t)n_part * 8) + cmc_align((size_t)nb * CMC_WORDS * 4) + 4 * cmc_align((size_t)nb * 4) +
                 cmc_align((size_t)nb * 8) + cmc_align((size_t)nb * 12 * 8) + 256;
    return (long long)tot;
}

extern "C" int hs_canting_moments_mc(hs_ctx *ctx, int n_cant, const double *d_sigma_deg, double ele_deg, int n_part,
                                     const unsigned long long *h_pcg_state, double slack, void *d_scratch, long long scratch_bytes,
                                     double *d_out, void *stream)
{
    if (!ctx) return -2;
    if (n_cant <= 0) return 0;
    if (n_part < CMC_B) { hs_set_err(ctx, "hs_canting_moments_mc: n_part must be at least 1024"); return -3; }
    if (!h_pcg_state || !d_scratch || scratch_bytes < hs_canting_mc_scratch_bytes(n_cant, n_part, slack)) {
        hs_set_err(ctx, "hs_canting_moments_mc: generator state missing or scratch smaller than hs_canting_mc_scratch_bytes"); return -3;
    }
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    hs::CmcArgs a;
    a.g0.s = hs::u128(h_pcg_state[0], h_pcg_state[1]);
    a.g0.inc = hs::u128(h_pcg_state[2], h_pcg_state[3]);
    a.n_part = n_part; a.n_cant = n_cant; a.nblocks = cmc_nblocks(n_cant, n_part, slack);
    a.sigma = d_sigma_deg;
    const double thet0 = 3.141592653589793 / 2.0 - ele_deg * 3.141592653589793 / 180.0;
    a.sin_t0 = sin(thet0); a.cos_t0 = cos(thet0);
    char *p = (char *)d_scratch;
    const long long nb = a.nblocks;
    a.ca = (double *)p; p += cmc_align((size_t)n_part * 8);
    a.sa = (double *)p; p += cmc_align((size_t)n_part * 8);
    a.bitmap = (unsigned *)p; p += cmc_align((size_t)nb * CMC_WORDS * 4);
    a.count0 = (int *)p; p += cmc_align((size_t)nb * 4);
    a.exit0 = (int *)p; p += cmc_align((size_t)nb * 4);
    a.entry = (int *)p; p += cmc_align((size_t)nb * 4);
    a.cnt = (int *)p; p += cmc_align((size_t)nb * 4);
    a.first = (long long *)p; p += cmc_align((size_t)nb * 8);
    a.partial = (double *)p; p += cmc_align((size_t)nb * 12 * 8);
    a.flag = (int *)p;
    CK(cudaMemsetAsync(a.flag, 0, sizeof(int), st));
    const unsigned gb = (unsigned)((nb + 127) / 128);
    hs::k_cmc_alpha<<<(n_part + 127) / 128, 128, 0, st>>>(a);
    hs::k_cmc_walk<<<gb, 128, 0, st>>>(a);
    hs::k_cmc_fix<<<gb, 128, 0, st>>>(a);
    hs::k_cmc_scan<<<1, 1024, 0, st>>>(a);
    hs::k_cmc_emit<<<gb, 128, 0, st>>>(a);
    hs::k_cmc_reduce<<<n_cant, 32, 0, st>>>(a, d_out);
    ctx->launches += 6;
    CK(cudaGetLastError());
    int flag = 0;
    CK(cudaMemcpyAsync(&flag, a.flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (flag == 2) return 2;   // not enough of the stream covered: call again with a larger slack
    if (flag) { hs_set_err(ctx, "hs_canting_moments_mc: sample orbits did not merge inside a block"); return -4; }
    return 0;
}

extern "C" int hs_canting_observables(hs_ctx *ctx, int n_tab, int n_D, const double *d_f, const double *d_mom, long long mom_stride_tab,
                                      long long mom_stride_D, const double *d_lam_tab, double kw_sqr, int n_dsd, const double *d_W,
                                      int n_extra, const double *d_extra, double *d_out, double *d_out_extra, void *stream)
{
    if (!ctx) return -2;
    if (n_tab <= 0 || n_D <= 0) return 0;
    if (n_extra < 0 || n_extra > CANT_NX || (n_extra > 0 && (!d_extra || !d_out_extra))) {
        hs_set_err(ctx, "hs_canting_observables: 0..4 extra columns, with their input and output arrays"); return -3;
    }
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    if (!d_W) {
        const long tot = (long)n_tab * n_D;
        hs::k_canting_sp<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(n_tab, n_D, d_f, d_mom, (long)mom_stride_tab, (long)mom_stride_D,
                                                                       d_lam_tab, kw_sqr, d_out);
    } else {
        if (n_dsd <= 0) return 0;
        dim3 grid((n_dsd + CANT_RB - 1) / CANT_RB, n_tab);
        hs::k_canting_psd<<<grid, CANT_RB * CANT_NC, 0, st>>>(n_dsd, n_D, n_tab, d_W, d_f, d_mom, (long)mom_stride_tab, (long)mom_stride_D,
                                                             d_lam_tab, kw_sqr, n_extra, d_extra, d_out, d_out_extra);
    }
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int hs_canting_mixture(hs_ctx *ctx, long long n, const double *d_a, const double *d_b, double *d_out, void *stream)
{
    if (!ctx) return -2;
    if (n <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    hs::k_canting_mixture<<<(unsigned)((n + 127) / 128), 128, 0, st>>>((long)n, d_a, d_b, d_out);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

// FP64 FMA peak: 8 independent DFMA chains per thread, enough CTAs to fill every SM.
__global__ void k_dfma_peak(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1e-9, a2 = a0 + 2e-9, a3 = a0 + 3e-9;
    double a4 = a0 + 4e-9, a5 = a0 + 5e-9, a6 = a0 + 6e-9, a7 = a0 + 7e-9;
    const double b = 1.0000001, c = 1e-12;
    for (int i = 0; i < iters; i++) {
        a0 = a0 * b + c; a1 = a1 * b + c; a2 = a2 * b + c; a3 = a3 * b + c;
        a4 = a4 * b + c; a5 = a5 * b + c; a6 = a6 * b + c; a7 = a7 * b + c;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" int hs_fp64_peak(hs_ctx *ctx, double *tflops, void *stream)
{
    if (!ctx) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    const int threads = 256, blocks = ctx->sm_count * 8, iters = 20000;
    double *d_out = nullptr;
    CK(cudaMalloc(&d_out, (size_t)threads * blocks * sizeof(double)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(e0, st));
        k_dfma_peak<<<blocks, threads, 0, st>>>(d_out, iters);
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        double fl = 2.0 * 8.0 * (double)iters * threads * blocks;
        double tf = fl / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    ctx->launches += 5;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
    *tflops = best;
    return 0;
}


// FP64 tensor-core (DMMA m8n8k4) peak: 8 independent accumulator pairs per warp.
__global__ void __launch_bounds__(256) k_dmma_peak(double *out, int iters)
{
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) hs::st_dmma(c + 2 * i, a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" int hs_fp64_tensor_peak(hs_ctx *ctx, double *tflops, void *stream)
{
    if (!ctx) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    const int threads = 256, blocks = ctx->sm_count * 4, iters = 20000;
    double *d_out = nullptr;
    CK(cudaMalloc(&d_out, (size_t)threads * blocks * sizeof(double)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(e0, st));
        k_dmma_peak<<<blocks, threads, 0, st>>>(d_out, iters);
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        // per warp and iteration: 8 DMMA x (8 x 8 x 4) FMA
        double fl = 2.0 * 8.0 * 256.0 * (double)iters * (threads / 32) * blocks;
        double tf = fl / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    ctx->launches += 5;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
    *tflops = best;
    return 0;
}

extern "C" int hs_tm2l_batch(hs_ctx *ctx, int n, const double *d_dpart, const double *d_dcore, const double *d_aovrb,
                             const double *d_aovrb2, const double *d_eor, const double *d_eoi, const double *d_eir,
                             const double *d_eii, double wavelength_cm, double *d_amp, int *d_status, void *stream)
{
    if (!ctx) return -2;
    if (n <= 0) return 0;
    if (!(wavelength_cm > 0.0)) { hs_set_err(ctx, "wavelength must be positive"); return -3; }
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    hs::T2Args a;
    a.n = n; a.dpart = d_dpart; a.dcore = d_dcore; a.aovrb = d_aovrb; a.aovrb2 = d_aovrb2;
    a.eor = d_eor; a.eoi = d_eoi; a.eir = d_eir; a.eii = d_eii; a.alamd = wavelength_cm;
    a.amp = d_amp; a.status = d_status; a.prof = nullptr;
    const bool prof = getenv("HYDROTM_PROF") != nullptr;
    if (prof) {
        if (!ctx->d_prof) CK(cudaMalloc(&ctx->d_prof, 24 * sizeof(long long)));
        CK(cudaMemsetAsync(ctx->d_prof, 0, 24 * sizeof(long long), st));
        a.prof = (unsigned long long *)ctx->d_prof;
    }
    size_t smem = (size_t)hs::t2_smem_doubles(2) * sizeof(double);
    CK(cudaFuncSetAttribute(hs::k_tm2l, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaMemsetAsync(ctx->d_queue + 8, 0, sizeof(int), st));
    int grid = std::min(n, ctx->sm_count);
    static const int split = getenv("HYDROTM_TM2L_SPLIT") ? std::max(1, std::min(4, atoi(getenv("HYDROTM_TM2L_SPLIT")))) : 1;
    hs::k_tm2l<<<grid, T2_ASM_THREADS + T2_SOLVE_THREADS, smem, st>>>(a, ctx->d_queue + 8, split == 3 ? 2 : split);
    ctx->launches++;
    CK(cudaGetLastError());
    if (prof) {
        unsigned long long h[24];
        CK(cudaStreamSynchronize(st));
        CK(cudaMemcpy(h, ctx->d_prof, sizeof(h), cudaMemcpyDeviceToHost));
        const char *nm[] = {"tables", "assemble", "asm wait", "chain", "chain wait", "solve1", "products", "solve2", "addprc"};
        const double np_ = (double)std::max<unsigned long long>(1, h[hs::T2P_PARTICLES]);
        fprintf(stderr, "[hydrotm-tm2l] %llu particles, kcycles per particle:", h[hs::T2P_PARTICLES]);
        for (int q = 0; q < hs::T2P_PARTICLES; q++) fprintf(stderr, " %s %.1f", nm[q], h[q] / np_ / 1e3);
        fprintf(stderr, "\n");
    }
    return 0;
}

extern "C" int hs_scatter_batch(hs_ctx *ctx, int n, const double *d_tarena, const long long *d_toff, const int *d_nmax,
                                const int *d_status, const double *d_lam, int n_geom, const double *h_geom, int n_orient,
                                const double *d_alpha, const double *d_beta, const double *d_weight, double scale,
                                double *d_S, double *d_Z, double *d_xs, void *stream)
{
    if (!ctx) return -2;
    if (n <= 0 || n_geom <= 0 || n_orient <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    ScPrep SP;
    { int rc = scatter_prepare(ctx, n, d_tarena, d_toff, d_nmax, d_status, d_lam, n_geom, h_geom, n_orient, d_alpha, d_beta, d_weight,
                               scale, d_S, d_Z, d_xs, st, SP); if (rc) return rc; }
    hs::ScArgs a = SP.a;
    const int n_inc = SP.n_inc;
    const int *dib = SP.inc_of_geom - 130;
    double *xs_inc = SP.xs_inc;
    // one thread per orientation (up to 64 per pass); dynamic smem = T stage (2*nmax^2 complex) + the per-orientation
    // reduction rows.  Large batches are ordered heaviest-first on the device and launched as three NMAX classes on
    // concurrent streams, each with a T stage sized for its own largest NMAX (more CTAs per SM for the small particles).
    const int threads = SP.threads;
    auto smem_for = [&](int nb) { return sc_smem(threads, nb); };
    // lanes per orientation of the three NMAX classes (20, NPN1], (10, 20], [0, 10]: HYDROTM_SC_SPLIT="h,m,s"
    int split[3] = {2, 1, 1};
    if (const char *e = getenv("HYDROTM_SC_SPLIT")) sscanf(e, "%d,%d,%d", &split[0], &split[1], &split[2]);
    a.order = nullptr;
    if (n >= 1024) {
        if (n > ctx->order_cap) {
            if (ctx->d_order) cudaFree(ctx->d_order);
            ctx->d_order = nullptr; ctx->order_cap = 0;
            CK(cudaMalloc(&ctx->d_order, (size_t)n * sizeof(int)));
            ctx->order_cap = n;
        }
        if (ensure_pin(ctx, (HS_NPN1 + 2) * sizeof(int))) return -1;
        int *hist = ctx->h_pin;
        hs::k_order_by_nmax<<<1, 1024, 0, st>>>(n, d_nmax, d_status, d_toff, ctx->d_order, hist);
        ctx->launches++;
        CK(cudaStreamSynchronize(st));
        const int bounds[3] = {HS_NPN1, 20, 10};  // classes (20, NPN1], (10, 20], [0, 10] in the sorted (descending) order
        int start = 0;
        size_t max_smem = 0;
        int cls_start[3], cls_cnt[3], cls_bound[3];
        for (int c = 0; c < 3; c++) {
            const int lo = c < 2 ? bounds[c + 1] + 1 : 0, hi = bounds[c];
            int cnt = 0, top = 0;
            for (int v = hi; v >= lo; v--) { cnt += hist[v]; if (hist[v] && !top) top = v; }
            cls_start[c] = start; cls_cnt[c] = cnt; cls_bound[c] = std::max(top, 1);
            start += cnt;
            if (cnt) max_smem = std::max(max_smem, smem_for(cls_bound[c]));
        }
        if (max_smem > (size_t)ctx->smem_optin) { hs_set_err(ctx, "scatter: nmax too large for the shared-memory T stage"); return -4; }
        CK(sc_set_smem((int)max_smem));
        { int top = 1; for (int c = 0; c < 3; c++) if (cls_cnt[c]) top = std::max(top, cls_bound[c]);
          int rc = scatter_build_wig(ctx, a, top, st); if (rc) return rc; }
        CK(cudaEventRecord(ctx->ev_start, st));
        for (int c = 0; c < 3; c++) {
            if (!cls_cnt[c]) continue;
            cudaStream_t bs = ctx->bstream[c];
            CK(cudaStreamWaitEvent(bs, ctx->ev_start, 0));
            hs::ScArgs ac = a;
            ac.order = ctx->d_order + cls_start[c];
            ac.nmax_bound = cls_bound[c];
            CK(sc_launch(split[c], cls_cnt[c], threads, smem_for(cls_bound[c]), bs, ac));
            ctx->launches++;
            CK(cudaEventRecord(ctx->ev_done[c], bs));
            CK(cudaStreamWaitEvent(st, ctx->ev_done[c], 0));
        }
    } else {
        int nmax_bound = ctx->last_nmax_max > 0 ? ctx->last_nmax_max : HS_NPN1;
        a.nmax_bound = nmax_bound;
        size_t smem = smem_for(nmax_bound);
        if (smem > (size_t)ctx->smem_optin) { hs_set_err(ctx, "scatter: nmax too large for the shared-memory T stage"); return -4; }
        CK(sc_set_smem((int)smem));
        { int rc = scatter_build_wig(ctx, a, nmax_bound, st); if (rc) return rc; }
        CK(sc_launch(nmax_bound > 20 ? split[0] : (nmax_bound > 10 ? split[1] : split[2]), n, threads, smem, st, a));
        ctx->launches++;
    }
    if (d_xs) {
        long tot = (long)n * n_geom * 2;
        k_xs_expand<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(n, n_geom, n_inc, dib + 130, xs_inc, d_xs);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    return 0;
}
