/*
 * hydrotm.h -- C ABI of libhydrotm.so, the B200 (sm_100a) T-matrix scattering engine.
 *
 * This is the drop-in boundary for hydroscatt's numerical back-ends.  Every entry point is
 * batched, takes plain device pointers + sizes + a CUDA stream (as void*), returns an int status
 * (0 = ok, <0 = call-level error, text via hs_last_error) and never throws.  Per-particle outcomes
 * are reported in a device status array (HS_ST_*).
 *
 * What each entry point replaces in the reference (paths relative to /root/reference/hydroscatt):
 *
 *   hs_calctmat_batch   pytmatrix's f2py routine `calctmat(axi,rat,lam,mrr,mri,eps,np,ddelt,ndgs)->nmax`
 *                       reached from Scatterer._init_tmatrix; call sites scattering.py:478-525,
 *                       scattering_rain.py:183-191 (one call per diameter).  SURVEY.md rows P2-P7.
 *   hs_calcampl_batch   pytmatrix's f2py routine `calcampl(nmax,lam,thet0,thet,phi0,phi,alpha,beta)->(S,Z)`
 *                       + orientation.orient_averaged_fixed (50 calls per particle and geometry);
 *                       call sites scattering.py:478-503 via radar.* / scatter.*.  Rows P8, P9.
 *   hs_xsect_batch      scatter.sca_xsect / scatter.ext_xsect (scipy dblquad over calcampl,
 *                       scattering.py:478-486, PSDIntegrator.init_scatter_table(angular_integration=True),
 *                       scattering_rain.py:282-283).  Row P12.
 *   hs_psd_integrate    PSDIntegrator.get_SZ / get_angular_integrated (trapz over D, row P10) and the
 *                       rectangle-rule sums of scattering.py:312-448 (row S5).
 *   hs_observables      radar.refl/Zdr/ldr/rho_hv/delta_hv/Kdp/Ai (row P12; scattering.py:530-587).
 *   hs_psd_observables  the two above fused: compute_scattering_psd_tm for every (DSD, table) in one kernel.
 *   hs_tm2l_batch       f_2l_tmatrix.tm2l.tmatrix(2) = hydroscatt/f_2l_tmatrix/tmatrix_py.f:1-473
 *                       (two-layer EBCM, rows F1-F11), called from scattering.py:32-67.
 *
 * T-matrix storage ("packed parity layout", complex128 interleaved):
 *   for m = 0..nmax: nmin = max(m,1), NM = nmax-nmin+1; two NM x NM matrices (parity class c=0,1),
 *   element [c][k'][k] (k = n-nmin scattered index, k' = n'-nmin incident index) holds
 *   T^{tau tau'}_{m n n'} with tau = 1 if (n+1)%2==c else 2, tau' = 1 if (n'+1)%2==c else 2.
 *   (All other T elements of a mirror-symmetric particle are exactly zero.)
 *   Mode m starts at complex offset 2*sum_{m'<m} NM(m')^2; a particle occupies hs_tmat_size(nmax).
 */
#ifndef HYDROTM_H
#define HYDROTM_H

#ifdef __cplusplus
extern "C" {
#endif

#define HS_NPN1 100   /* NMAX limit (upstream ampld.par.f NPN1) */
#define HS_NPNG1 500  /* NGAUSS limit (upstream NPNG1) */

/* per-particle status codes */
#define HS_RC_ARENA_FULL 1  /* return value of hs_calctmat_batch: some particles have HS_ST_ARENA_FULL; *d_arena_top = entries needed */
#define HS_ST_OK 0
#define HS_ST_NMAX_LIMIT 1    /* NMAX reached HS_NPN1 without convergence (upstream: print + STOP) */
#define HS_ST_NGAUSS_LIMIT 2  /* NGAUSS reached HS_NPNG1 (upstream: warning, continues) */
#define HS_ST_SINGULAR 3      /* zero pivot in the Q solve */
#define HS_ST_ARENA_FULL 4    /* caller's T arena too small for this particle */
#define HS_ST_BAD_INPUT 5
#define HS_ST_INTERNAL_ESCALATE 100 /* never returned to the caller */
#define HS_ST_INTERNAL_GAUSS 101    /* never returned to the caller */

#define HS_SHAPE_SPHEROID (-1)
#define HS_SHAPE_CYLINDER (-2) /* finite circular cylinder, axis_ratio = diameter / length (Mishchenko NP = -2) */

typedef struct hs_ctx hs_ctx;

/* Library version / build info; safe to call without a GPU. */
const char *hs_version(void);
/* complex128 entries a particle with truncation nmax occupies in the T arena. */
long long hs_tmat_size(int nmax);
/* complex offset of mode m inside a particle's packed T-matrix. */
long long hs_tmat_mode_offset(int nmax, int m);
/* upstream's starting truncation max(4, int(x + 4.05 x**0.333333)), x = 2 pi r / lambda */
int hs_nmax_start(double r_ev, double wavelength);

/* Create / destroy an engine context on CUDA device `device`.  Returns 0 or a negative error. */
int hs_ctx_create(hs_ctx **ctx, int device);
void hs_ctx_destroy(hs_ctx *ctx);
const char *hs_last_error(const hs_ctx *ctx);
/* kernels launched by this context so far (for bench.py's gpu_launches) */
long long hs_launch_count(const hs_ctx *ctx);

/*
 * Single-layer T-matrix for n spheroids (converged NMAX/NGAUSS per particle + all azimuthal modes).
 *   d_axi,d_lam,d_mr,d_mi,d_eps : device [n]  radius, wavelength (same unit), Re m, Im m, axis ratio
 *                                 (horizontal/vertical, >1 oblate)
 *   rat      : 1.0 = d_axi is the equal-volume radius; other = equal-surface-area radius (SAREA)
 *   np       : HS_SHAPE_SPHEROID or HS_SHAPE_CYLINDER
 *   ddelt    : convergence tolerance as the caller states it; rescale_ddelt != 0 applies
 *              Mishchenko's DDELT = 0.1*DDELT before use
 *   d_tarena : device buffer of arena_cap complex128; d_arena_top device counter (complex entries
 *              used so far; the call appends).  d_toff[n] receives each particle's offset.
 *   d_nmax,d_ngauss,d_ntrials,d_status : device [n] outputs.
 * The call enqueues on `stream` and synchronises it before returning.
 */
int hs_calctmat_batch(hs_ctx *ctx, int n, const double *d_axi, const double *d_lam, const double *d_mr,
                      const double *d_mi, const double *d_eps, double rat, int np, double ddelt,
                      int ndgs, int rescale_ddelt, double *d_tarena, long long arena_cap,
                      unsigned long long *d_arena_top, long long *d_toff, int *d_nmax, int *d_ngauss,
                      int *d_ntrials, int *d_status, void *stream);

/*
 * Amplitude (S, 2x2 complex) and phase (Z, 4x4 real) matrices.
 *   d_geom   : device [n_geom][4]  thet0, thet, phi0, phi (degrees), shared by all particles
 *   d_alpha,d_beta,d_weight : device [n_orient] Euler angles (degrees) and quadrature weights
 *   reduce=1 : d_S [n][n_geom][4] complex, d_Z [n][n_geom][16] = scale * sum_o weight[o]*(S,Z)(o)
 *   reduce=0 : d_S [n][n_geom][n_orient][4], d_Z [n][n_geom][n_orient][16] (weights ignored)
 * Particles whose d_status != 0 get NaN.  Asynchronous on `stream`.
 */
int hs_calcampl_batch(hs_ctx *ctx, int n, const double *d_tarena, const long long *d_toff,
                      const int *d_nmax, const int *d_status, const double *d_lam, int n_geom,
                      const double *d_geom, int n_orient, const double *d_alpha, const double *d_beta,
                      const double *d_weight, double scale, int reduce, double *d_S, double *d_Z,
                      void *stream);

/*
 * Orientation-averaged scattering cross sections (h- and v-polarised incidence) from the T-matrix expansion
 * coefficients; equals the sphere integral of Z11 -/+ Z12 that scatter.sca_xsect evaluates with dblquad.
 *   d_inc : device [n_inc][2] thet0, phi0 (degrees);  d_out : device [n][n_inc][2] (sca_h, sca_v)
 *   result = scale * sum_o weight[o] * C(o).  Asynchronous on `stream`.
 */
int hs_xsect_batch(hs_ctx *ctx, int n, const double *d_tarena, const long long *d_toff, const int *d_nmax,
                   const int *d_status, const double *d_lam, int n_inc, const double *d_inc, int n_orient,
                   const double *d_alpha, const double *d_beta, const double *d_weight, double scale,
                   double *d_out, void *stream);

/*
 * Fused scatter-table evaluation: orientation-averaged S and Z for n_geom geometries AND the h/v scattering cross
 * sections of each geometry's incidence direction, from ONE sweep over the T-matrix per (orientation, incidence
 * direction) (hs_calcampl_batch + hs_xsect_batch in a single kernel; see csrc/hs_scatter.cuh).
 *   h_geom : HOST [n_geom][4] thet0, thet, phi0, phi (degrees); geometries sharing (thet0, phi0) share the sweep
 *   d_S [n][n_geom][4] complex, d_Z [n][n_geom][16], d_xs [n][n_geom][2] (sca_xsect h, v; may be NULL)
 *   results = scale * sum_o weight[o] * value(o).  Asynchronous on `stream`.
 */
int hs_scatter_batch(hs_ctx *ctx, int n, const double *d_tarena, const long long *d_toff, const int *d_nmax,
                     const int *d_status, const double *d_lam, int n_geom, const double *h_geom, int n_orient,
                     const double *d_alpha, const double *d_beta, const double *d_weight, double scale, double *d_S,
                     double *d_Z, double *d_xs, void *stream);

/*
 * PSD weights W[d][i] = wq[i] * N_d(D[i]) for n_dsd size distributions (pytmatrix.psd classes, row P11).
 *   kind 0: GammaPSD (p0=D0, p1=Nw, p2=mu)   1: ExponentialPSD (p0=N0, p1=Lambda)
 *   kind 2: UnnormalizedGammaPSD (p0=N0, p1=Lambda, p2=mu);   d_dmax[d] = D_max cut-off
 *   d_wq[n_D]: quadrature weights over D (trapezoid for PSDIntegrator, D[i]-D[i-1] for scattering.py:312-448).
 */
int hs_psd_weights(hs_ctx *ctx, int kind, int n_dsd, const double *d_p0, const double *d_p1, const double *d_p2,
                   const double *d_dmax, int n_D, const double *d_D, const double *d_wq, double *d_W,
                   void *stream);

/* PSD integration as an FP64 GEMM: out[n_dsd][ncol] = W[n_dsd][n_D] * Tt[n_D][ncol]. */
int hs_psd_integrate(hs_ctx *ctx, int n_dsd, int n_D, int ncol, const double *d_W, const double *d_Tt,
                     double *d_out, void *stream);

/*
 * hs_calctmat_batch + hs_scatter_batch in one call (what scattering_rain.py:183-191 + PSDIntegrator.init_scatter_table do per
 * diameter: calctmat, then 2 x 50 calcampl and the cross sections).  With HYDROTM_OVERLAP_SCATTER=1 particles whose T-matrix
 * is complete are scattered right away on a side stream while the others are still in their convergence rounds (measured
 * slower than back to back on B200, hence opt-in).  Arguments = those of the two calls;
 * returns like hs_calctmat_batch (HS_RC_ARENA_FULL: retry with a larger arena, outputs undefined).
 */
int hs_calctmat_scatter_batch(hs_ctx *ctx, int n, const double *d_axi, const double *d_lam, const double *d_mr, const double *d_mi,
                              const double *d_eps, double rat, int np, double ddelt, int ndgs, int rescale_ddelt, double *d_tarena,
                              long long arena_cap, unsigned long long *d_arena_top, long long *d_toff, int *d_nmax, int *d_ngauss,
                              int *d_ntrials, int *d_status, int n_geom, const double *h_geom, int n_orient, const double *d_alpha,
                              const double *d_beta, const double *d_weight, double scale, double *d_S, double *d_Z, double *d_xs,
                              void *stream);

/*
 * Fused PSD integration + observables (what compute_scattering_psd_tm, scattering.py:530-587, does one DSD at a time
 * through PSDIntegrator.get_SZ + radar.*): obs[d][t][15] (columns as hs_observables) for n_dsd weight rows
 * d_W [n_dsd][n_D] and n_tab tables of records d_rows [n_tab][n_D][56] (layout: hydroscatt_b200/tables.py), without
 * materialising the integrated [n_dsd][n_tab][56] array.  hpol_quirk != 0: ext_xsect_v := ext_xsect_h (upstream
 * scatter.ext_xsect ignores h_pol once an angular table exists).  want_sca == 0: sca_xsect columns are NaN.
 */
int hs_psd_observables(hs_ctx *ctx, int n_dsd, int n_D, int n_tab, const double *d_W, const double *d_rows,
                       const double *d_lam_tab, double kw_sqr, int hpol_quirk, int want_sca, double *d_obs, void *stream);

/*
 * Same, writing only the n_sel (1..15) observable columns the caller names, in the caller's order: d_obs [n_dsd][n_tab][n_sel],
 * h_sel = HOST array of column indices into the 15-column list of hs_observables (e.g. {2, 4, 11, 7, 12, 14} = Zh, Zdr, Kdp,
 * rhohv, Ah, Adp, the columns scattering_rain.py:296-316 keeps per DSD).  The scattering cross sections (columns 0, 1) are
 * integrated only when selected.
 */
int hs_psd_observables_sel(hs_ctx *ctx, int n_dsd, int n_D, int n_tab, const double *d_W, const double *d_rows,
                           const double *d_lam_tab, double kw_sqr, int hpol_quirk, int n_sel, const int *h_sel, double *d_obs,
                           void *stream);

/*
 * Table records (56 doubles per particle, layout in hydroscatt_b200/tables.py: what PSDIntegrator.init_scatter_table keeps per
 * diameter as _S_table / _Z_table / _angular_table) from the outputs of hs_scatter_batch: d_S [n][n_geom][2][2] complex,
 * d_Z [n][n_geom][4][4], d_xs [n][n_geom][2], d_lam [n].  g_back / g_forw: index of the back / forward geometry; f_back /
 * f_forw: index of the geometry whose S is the forward amplitude at that geometry's incidence direction (ext_xsect by the
 * optical theorem, scatter.ext_xsect).  d_rows: [n][56].
 */
int hs_pack_rows(hs_ctx *ctx, int n, int n_geom, const double *d_S, const double *d_Z, const double *d_xs, const double *d_lam,
                 int g_back, int g_forw, int f_back, int f_forw, double *d_rows, void *stream);

/*
 * d_dst[i][0..ncol) = d_src[d_index[i]][0..ncol) for i < n_out (records of ncol doubles; d_index: device int64, negative = leave
 * d_dst[i] untouched).  Used by the multi-GPU table build to put all-gathered shard records back into table order.
 */
int hs_gather_rows(hs_ctx *ctx, long long n_out, int ncol, const double *d_src, const long long *d_index, double *d_dst,
                   void *stream);

/*
 * Radar observables per row (radar.refl/Zdr/ldr/rho_hv/delta_hv/Kdp/Ai, scatter.sca_xsect/ext_xsect as combined
 * by scattering.py:451-587).  d_in: [n][stride]; column offsets: off_back / off_forw -> S(8 doubles) Z(16 doubles)
 * of the back / forward geometry, off_ext -> ext_xsect_h, ext_xsect_v (or -1: optical theorem on the forward S),
 * off_sca -> sca_xsect_h, sca_xsect_v (or -1).  d_lam[i*lam_stride] = wavelength of row i.
 * d_out: [n][15] = sca_xsect_h, sca_xsect_v, refl_h, refl_v, zdr, ldr_h, ldr_v (dB), rho_hv, delta_hv (deg),
 *                  ext_xsect_h, ext_xsect_v (dB), kdp (deg/km), A_h, A_v, Adp (dB/km)
 */
int hs_observables(hs_ctx *ctx, int n, int stride, const double *d_in, int off_back, int off_forw, int off_ext,
                   int off_sca, const double *d_lam, int lam_stride, double kw_sqr, double *d_out, void *stream);

/*
 * Canting-angle ("angular moments") observables: batched scattering.compute_scattering_canting_sp (scattering.py:182-309)
 * and compute_scattering_canting_psd (scattering.py:312-448; looped 9 300x at scattering_rain.py:295-316 and 11 011x at
 * scattering_hail.py:298-372).  SURVEY.md rows S4, S5.
 *   d_f    : device [n_tab][n_D][8]  fv180, fh180, fv0, fh0 (re, im; metres) per table (species / band / temperature)
 *   d_mom  : device angular moments a1, a2, a3, a4, a5, a7 of the canting distribution, 6 contiguous doubles at
 *            d_mom[t * mom_stride_tab + i * mom_stride_D] (strides in doubles; 0 = shared by all tables / diameters)
 *   d_lam_tab : device [n_tab] wavelength [mm];  kw_sqr: dielectric factor of the reflectivity
 *   d_W    : device [n_dsd][n_D] = psd_vals * delta_d (rectangle rule), or NULL for the single-particle form
 *   d_extra: device [n_D][n_extra] (n_extra <= 4) extra per-diameter columns summed with the same weights (e.g. D^3 and
 *            D^3 v(D) for precip.compute_lwc / compute_rainfall_rate), or NULL;  d_out_extra: device [n_dsd][n_extra]
 *   d_out  : d_W == NULL: [n_tab][n_D][15];  else [n_dsd][n_tab][15].  Columns (the reference's DataFrame order):
 *            sca_xsect_h, sca_xsect_v, ext_xsect_h, ext_xsect_v, ldr_h, ldr_v, refl_h, refl_v, zdr [dB], rho_hv,
 *            delta_hv [deg], kdp [deg/km], A_h, A_v, Adp [dB/km].
 * Asynchronous on `stream`.
 */
int hs_canting_observables(hs_ctx *ctx, int n_tab, int n_D, const double *d_f, const double *d_mom, long long mom_stride_tab,
                           long long mom_stride_D, const double *d_lam_tab, double kw_sqr, int n_dsd, const double *d_W,
                           int n_extra, const double *d_extra, double *d_out, double *d_out_extra, void *stream);

/*
 * Monte-Carlo angular moments of the canting-angle distribution, scattering.compute_angular_moments (scattering.py:70-133, SURVEY.md
 * row S2 / 8(f)3), on the SAME random stream as the reference: numpy.random.default_rng(seed) -> one uniform(0, 360) vector of
 * n_part alphas, then n_part normal(scale = sigma) betas per canting angle, in order (PCG64 + numpy's 256-layer ziggurat).
 *   d_sigma_deg : device [n_cant] canting angles (deg);  ele_deg: elevation angle;  n_part >= 1024 (reference default 500 000)
 *   h_pcg_state : HOST {state_hi, state_lo, inc_hi, inc_lo} of the PCG64 generator before its first draw
 *                 (numpy.random.default_rng(seed).bit_generator.state["state"])
 *   slack       : stream positions provided per sample (>= 1; the ziggurat consumes ~1.013 draws per normal: pass 1.03)
 *   d_scratch   : device scratch of hs_canting_mc_scratch_bytes(n_cant, n_part, slack) bytes
 *   d_out       : device [n_cant][6] = a1, a2, a3, a4, a5, a7 (means over the n_part samples)
 * Returns 0, or 2 if `slack` was too small (call again with a larger one); synchronises `stream`.
 */
long long hs_canting_mc_scratch_bytes(int n_cant, int n_part, double slack);
int hs_canting_moments_mc(hs_ctx *ctx, int n_cant, const double *d_sigma_deg, double ele_deg, int n_part,
                          const unsigned long long *h_pcg_state, double slack, void *d_scratch, long long scratch_bytes,
                          double *d_out, void *stream);

/*
 * scattering.compute_scattering_mixture (scattering.py:590-673, row S7) on n rows of the 15 canting observables
 * (incoherent sum of two species); the cross-section columns 0..3, which the reference's mixture does not define, are NaN.
 */
int hs_canting_mixture(hs_ctx *ctx, long long n, const double *d_a, const double *d_b, double *d_out, void *stream);

/*
 * Two-layer (coated spheroid) T-matrix amplitudes for n particles at one wavelength: the batched equivalent of
 * tm2l.tmatrix(2) = tmatrix_py.f (fixed nrank 17, 7 azimuthal modes, 31-node Patterson rule; rows F1-F11).
 *   d_dpart,d_dcore : device [n] outer / inner equal-volume diameter [cm]   (fort.20 columns 1-2, tmatrix_py.f:580)
 *   d_aovrb,d_aovrb2: device [n] outer / inner axis ratio a/b (<1 oblate)   (columns 3-4)
 *   d_eor,d_eoi,d_eir,d_eii : device [n] Re/Im permittivity of the outer layer, then the inner body (columns 5-8)
 *   wavelength_cm   : namelist value palamd of spongyice.inp
 *   d_amp : device [n][8] = borrp borip borrpe boripe forrp forip forpe foripe [m] (tmatrix_py.f:2515-2520,
 *           full double precision; the e15.7 rounding belongs to the file layer);  d_status : device [n].
 * Asynchronous on `stream`.
 */
int hs_tm2l_batch(hs_ctx *ctx, int n, const double *d_dpart, const double *d_dcore, const double *d_aovrb,
                  const double *d_aovrb2, const double *d_eor, const double *d_eoi, const double *d_eir,
                  const double *d_eii, double wavelength_cm, double *d_amp, int *d_status, void *stream);

/* Measured FP64 FMA peak of the device (dependent-chain-free DFMA loop), TFLOP/s; used as roofline denominator. */
int hs_fp64_peak(hs_ctx *ctx, double *tflops, void *stream);
/* Measured FP64 tensor-core peak (DMMA m8n8k4, the instruction the Q / RgQ assembly runs on), TFLOP/s. */
int hs_fp64_tensor_peak(hs_ctx *ctx, double *tflops, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* HYDROTM_H */
