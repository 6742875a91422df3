// emu_driver.cpp -- TEST INFRASTRUCTURE ONLY: runs the device functions of hydroscatt_b200/csrc as
// single-lane host code (HS_HOST_EMU) so the kernel logic can be checked against the oracle on a box
// without a GPU.  Never loaded by the product.
#define HS_HOST_EMU 1
#include "../../hydroscatt_b200/csrc/hs_gauss.h"
#include "../../hydroscatt_b200/csrc/hs_tmat.cuh"
#include "../../hydroscatt_b200/csrc/hs_ampl.cuh"
#include <cstdlib>

extern "C" int emu_calctmat(int n, const double *axi, const double *lam, const double *mr, const double *mi,
                            const double *eps, double rat, double ddelt, int ndgs, int rescale, double *tarena,
                            long long cap, long long *toff, int *nmax, int *ngauss, int *ntrials, int *status,
                            int ws_cap, int gmax)
{
    size_t tot = (size_t)gmax * (gmax + 1) / 2;
    std::vector<double> hx(tot), hw(tot);
    for (int g = 1; g <= gmax; g++) hs::gauss_positive(2 * g, &hx[(size_t)g * (g - 1) / 2], &hw[(size_t)g * (g - 1) / 2]);
    const int T1 = HS_NPN1 + 1;
    std::vector<double> an(4 * T1, 0.0);
    double *dd_ = an.data() + T1;
    for (int k = 1; k <= HS_NPN1; k++) { an[k] = (double)(k * (k + 1)); dd_[k] = sqrt((double)(2 * k + 1) / (double)(k * (k + 1))); }
    { double cacc = 1.0; an[2 * T1] = 1.0; for (int i = 1; i <= HS_NPN1; i++) { cacc = cacc * sqrt((double)(2 * i - 1) / (double)(2 * i)); an[2 * T1 + i] = cacc; } }
    for (int k = 0; k <= HS_NPN1; k++) an[3 * T1 + k] = 1.0 / (double)(2 * k + 1);
    struct { double *p; double *data() { return p; } } dd{dd_};
    unsigned long long top = 0;
    int queue = 0;
    hs::TmArgs a;
    a.items = nullptr; a.n_items = n; a.queue = &queue;
    a.axi = axi; a.lam = lam; a.mr = mr; a.mi = mi; a.eps = eps;
    a.rat = rat; a.np = -1; a.ddelt = rescale ? 0.1 * ddelt : ddelt; a.ndgs = ndgs;
    a.gx = hx.data(); a.gw = hw.data(); a.gmax = gmax;
    a.tarena = tarena; a.arena_cap = cap; a.arena_top = &top;
    a.toff = toff; a.nmax = nmax; a.ngauss = ngauss; a.ntrials = ntrials; a.status = status;
    a.ws_cap = ws_cap; a.ws_global = nullptr; a.ws_fb = nullptr; a.fb_cap = 0; a.debug = 0; a.prof = nullptr; a.tile_min = getenv("HS_EMU_TILE_MIN") ? atoi(getenv("HS_EMU_TILE_MIN")) : 1000;
    std::vector<double> ws((size_t)ws_cap + 16);
    hs::Grp<true> grp; grp.tid = 0; grp.nt = 1;
    hs::Ctl ctl;
    for (int i = 0; i < n; i++) if (a.tile_min < 1000) hs::calctmat_particle<true, true>(a, i, ws.data(), nullptr, grp, &ctl, an.data(), dd.data());
        else hs::calctmat_particle<true, false>(a, i, ws.data(), nullptr, grp, &ctl, an.data(), dd.data());
    return 0;
}

extern "C" void emu_ampl(const double *t, int nmax, double lam, double thet0, double thet, double phi0, double phi,
                         double alpha, double beta, double *out24)
{
    std::vector<double> fn(HS_NPN1 + 1);
    for (int k = 1; k <= HS_NPN1; k++) fn[k] = sqrt((double)(2 * k + 1) / (double)(k * (k + 1)));
    hs::ampl_one(reinterpret_cast<const double2 *>(t), nmax, lam, thet0, thet, phi0, phi, alpha, beta, fn.data(), out24);
}

#include "../../hydroscatt_b200/csrc/hs_xsect.cuh"
extern "C" void emu_xsect(const double *t, int nmax, double lam, double thet0, double phi0, double alpha, double beta,
                          double *out2)
{
    std::vector<double> fn(HS_NPN1 + 1);
    for (int k = 1; k <= HS_NPN1; k++) fn[k] = sqrt((double)(2 * k + 1) / (double)(k * (k + 1)));
    hs::xsect_one(reinterpret_cast<const double2 *>(t), nmax, lam, thet0, phi0, alpha, beta, fn.data(), out2);
}

#include "../../hydroscatt_b200/csrc/hs_tm2l.cuh"
extern "C" void emu_tm2l(int n, const double *p8, double alamd, double *amp, int *status)
{
    std::vector<double> col[8];
    for (int c = 0; c < 8; c++) { col[c].resize(n); for (int i = 0; i < n; i++) col[c][i] = p8[8 * i + c]; }
    hs::T2Args a;
    a.n = n; a.dpart = col[0].data(); a.dcore = col[1].data(); a.aovrb = col[2].data(); a.aovrb2 = col[3].data();
    a.eor = col[4].data(); a.eoi = col[5].data(); a.eir = col[6].data(); a.eii = col[7].data();
    a.alamd = alamd; a.amp = amp; a.status = status;
    std::vector<double> smem(hs::t2_smem_doubles(1) + 16);
    hs::Ctl ctl;
    double cmx[sizeof(hs::T2Prep) / sizeof(double) + 2], acc[8];
    for (int i = 0; i < n; i++) { ctl.sing = 0; hs::tm2l_particle(a, i, smem.data(), &ctl, cmx, acc, 0, 1); }
}

#include "../../hydroscatt_b200/csrc/hs_scatter.cuh"
extern "C" void emu_scatter(const double *t, int nmax, double lam, double thet0, double phi0, int ng, const double *gsc,
                            double alpha, double beta, double *out, double *xs2)
{
    std::vector<double> fn(HS_NPN1 + 1);
    for (int k = 1; k <= HS_NPN1; k++) fn[k] = sqrt((double)(2 * k + 1) / (double)(k * (k + 1)));
    hs::scatter_one(reinterpret_cast<const double2 *>(t), nmax, lam, thet0, phi0, ng, gsc, alpha, beta, fn.data(), out, xs2);
}

// hs_gauss.h: the n-point rule (any n; cylinders use odd orders too) and the positive half of an even rule
extern "C" void emu_gauss_full(int n, double *z, double *w) { hs::gauss_full(n, z, w); }
extern "C" void emu_gauss_positive(int n, double *xp, double *wp) { hs::gauss_positive(n, xp, wp); }

// ---- device-driven rounds: controller + planner (hs_dvctl.h) on a table of trial results ----
// One particle: starting order s0, lower-bound guess lb; the trial results come from a table of n_tab known trials
// (NMAX, NGAUSS) -> (Qsca, Qext) -- the oracle's recorded sequence; a planned trial that is not in the table gets poison
// values, so a controller that consumed a trial beyond the reference's sequence cannot land on the reference's answer.
// out: nmax, ngauss, ntrials, status, done, rounds, planned items, consumed-unknown flag.
#include "../../hydroscatt_b200/csrc/hs_dvctl.h"
extern "C" void emu_dv_rounds(int s0, int lb, int ndgs, int spec, int spec0, int margin, double ddelt, int n_tab,
                              const int *tab_nmax, const int *tab_g, const double *tab_qsca, const double *tab_qext, int *out8)
{
    hs::DvP s;
    hs::dv_init(s, 0, s0, lb, spec0);
    int rounds = 0, planned = 0, unknown = 0;
    std::vector<hs::DvTrial> res;
    std::vector<int> known;
    while (!s.done && rounds < 2000) {
        res.clear(); known.clear();
        hs::dv_items_of(s, ndgs, spec, 1 << 30, 1 << 30, -1, nullptr, [&](int nmax, int g, int nmodes) {
            hs::DvTrial t; t.qsca = 1.2345e300; t.qext = -7.7e-300; t.t_off = nmodes > 1 ? 1000 + planned : -1; t.sing = 0; t.g = g;
            int k = 0;
            for (int i = 0; i < n_tab; i++) if (tab_nmax[i] == nmax && tab_g[i] == g) { t.qsca = tab_qsca[i]; t.qext = tab_qext[i]; k = 1; }
            res.push_back(t); known.push_back(k);
            planned++;
        });
        s.first_item = 0; s.n_items = (int)res.size();
        const int before = s.ntr;
        hs::dv_consume(s, ddelt, ndgs, margin, [&](int q) { return res[q]; });
        for (int q = 0; q < s.ntr - before; q++) if (!known[q]) unknown = 1;
        rounds++;
    }
    out8[0] = s.nmax; out8[1] = s.g; out8[2] = s.ntr; out8[3] = s.status; out8[4] = s.done; out8[5] = rounds; out8[6] = planned; out8[7] = unknown;
}
