"""CUDA back-end object behind the pytmatrix shim (numpy in / numpy out; everything numerical runs through the
C-ABI on the device).  Mirrors what pytmatrix's f2py module offered (calctmat / calcampl) plus the fused batched
calls the shim uses instead of Python loops."""
import os

import numpy as np
import torch

from . import engine as _engine

#: Mishchenko's ``DDELT = 0.1*DDELT`` rescale inside calctmat (hard part 1).  ON = original
#: ampld.lp.f behaviour.  Recorded in every result the table layer writes.
RESCALE_DDELT = os.environ.get("HYDROTM_RESCALE_DDELT", "1") not in ("0", "false", "False")


class ConvergenceError(RuntimeError):
    """Raised where upstream calctmat would print 'CONVERGENCE IS NOT OBTAINED' and STOP the interpreter."""


class TMHandle:
    """One particle's device-resident T-matrix."""

    def __init__(self, batch):
        self.batch = batch
        self.nmax_host = int(batch.nmax[0].item())
        self.ngauss_host = int(batch.ngauss[0].item())
        self.status_host = int(batch.status[0].item())
        self.toff_host = int(batch.toff[0].item())

    @property
    def nmax(self):
        return self.nmax_host


def trapz_weights(D):
    D = np.asarray(D, dtype=np.float64)
    w = np.zeros_like(D)
    if D.size > 1:
        d = np.diff(D)
        w[:-1] += 0.5 * d
        w[1:] += 0.5 * d
    return w


def rect_weights(D):
    """Rectangle rule of scattering_rain.py:271: delta_d = D - D_prev (first bin: D[0])."""
    D = np.asarray(D, dtype=np.float64)
    return D - np.concatenate([[0.0], D[:-1]])


class CudaBackend:
    def __init__(self, device=None):
        self.eng = _engine.default_engine(device)

    # ---- per-call equivalents of the f2py routines --------------------------------------------------------
    def calctmat(self, radius, radius_type, wavelength, m, axis_ratio, shape, ddelt, ndgs):
        tb = self.eng.calctmat([radius], [wavelength], [m.real], [m.imag], [axis_ratio], rat=float(radius_type),
                               shape=int(shape), ddelt=ddelt, ndgs=ndgs, rescale_ddelt=RESCALE_DDELT)
        h = TMHandle(tb)
        if h.status_host not in (_engine.ST_OK, _engine.ST_NGAUSS_LIMIT):
            raise ConvergenceError("calctmat: " + _engine.STATUS_TEXT.get(h.status_host, str(h.status_host)))
        if h.toff_host < 0:
            # NMAX*NDGS outran NPNG1 before the first NGAUSS trial: upstream prints and STOPs; there is no T-matrix
            raise ConvergenceError("calctmat: NGAUSS = NMAX*NDGS exceeds NPNG1=500 before convergence (no T-matrix)")
        return h

    def calcampl(self, h, thet0, thet, phi0, phi, alpha, beta):
        S, Z = h.batch.ampl([(thet0, thet, phi0, phi)], [alpha], [beta])
        return S[0, 0, 0].cpu().numpy(), Z[0, 0, 0].cpu().numpy()

    def orient_avg(self, h, geom4, alpha, beta, weight, scale):
        S, Z = h.batch.ampl([geom4], alpha, beta, weight=weight, scale=scale, reduce=True)
        return S[0, 0].cpu().numpy(), Z[0, 0].cpu().numpy()

    def xsect(self, h, inc, alpha, beta, weight, scale, want_asym=False):
        out = h.batch.xsect(inc, alpha, beta, weight, scale).cpu().numpy()
        res = {"sca_xsect_h": out[:, :, 0], "sca_xsect_v": out[:, :, 1]}
        if want_asym:
            a = self._asym(h.batch, inc, alpha, beta, weight, scale, h.nmax_host)
            res["asym_h"], res["asym_v"] = a[:, :, 0], a[:, :, 1]
        return res

    def _asym(self, tb, inc, alpha, beta, weight, scale, nmax_max):
        """Asymmetry parameter <cos Theta> for h / v incidence: [n, n_inc, 2].

        scatter.asym integrates I(theta, phi) cos(Theta) with dblquad; the integrand is a trigonometric polynomial of
        degree <= 2 NMAX + 1, so a product rule (NMAX+2 Gauss nodes in cos theta x 2 NMAX+4 equispaced phi) is exact.
        The orientation-averaged Z at all quadrature directions comes from the fused scatter kernel (64 geometries per
        launch, the incident-side T sweep shared by every direction of a launch pair)."""
        nt, nph = int(nmax_max) + 2, 2 * int(nmax_max) + 4
        x, w = np.polynomial.legendre.leggauss(nt)
        th = np.degrees(np.arccos(x))
        ph = (np.arange(nph) + 0.5) * 360.0 / nph
        TH, PH = np.meshgrid(th, ph, indexing="ij")
        WQ = np.repeat(w[:, None], nph, axis=1) * (2.0 * np.pi / nph)
        out = np.zeros((tb.n, len(inc), 2))
        for d, (t0, p0) in enumerate(inc):
            geoms = [(t0, a, p0, b) for a, b in zip(TH.reshape(-1), PH.reshape(-1))]
            Zs = []
            for c0 in range(0, len(geoms), 64):
                _, Z, _ = tb.scatter(geoms[c0:c0 + 64], alpha, beta, weight=weight, scale=scale, want_xs=False)
                Zs.append(Z.cpu().numpy())
            Z = np.concatenate(Zs, axis=1)                       # [n, ndir, 4, 4]
            cosT = (np.sin(np.radians(t0)) * np.sin(np.radians(TH)) * np.cos(np.radians(PH - p0)) +
                    np.cos(np.radians(t0)) * np.cos(np.radians(TH))).reshape(-1)
            wq = WQ.reshape(-1)
            for k, sgn in enumerate((-1.0, 1.0)):                # h: Z11 - Z12, v: Z11 + Z12
                I = Z[:, :, 0, 0] + sgn * Z[:, :, 0, 1]
                out[:, d, k] = (I * (wq * cosT)[None, :]).sum(1) / (I * wq[None, :]).sum(1)
        return out

    # ---- fused batched table ------------------------------------------------------------------------------
    def scatter_table(self, radii, radius_type, wavelength, m, axis_ratio, shape, ddelt, ndgs, geometries, alpha,
                      beta, weight, scale, angular):
        """S [n,G,2,2], Z [n,G,4,4] (+ sca/ext/asym tables [n,G]) for all diameters in three device calls."""
        m = np.asarray(m, dtype=np.complex128)
        tb = self.eng.calctmat(radii, wavelength, m.real, m.imag, axis_ratio, rat=float(radius_type),
                               shape=int(shape), ddelt=ddelt, ndgs=ndgs, rescale_ddelt=RESCALE_DDELT)
        st = tb.status.cpu().numpy()
        bad = np.nonzero((st != _engine.ST_OK) & (st != _engine.ST_NGAUSS_LIMIT))[0]
        if bad.size:
            raise ConvergenceError("calctmat: particle %d: %s" % (bad[0], _engine.STATUS_TEXT.get(int(st[bad[0]]))))
        empty = np.nonzero(tb.toff.cpu().numpy() < 0)[0]
        if empty.size:
            raise ConvergenceError("calctmat: particle %d: NGAUSS = NMAX*NDGS exceeds NPNG1=500 before convergence "
                                   "(no T-matrix)" % empty[0])
        g4 = [tuple(g[:4]) for g in geometries]
        res = {"nmax": tb.nmax.cpu().numpy(), "ngauss": tb.ngauss.cpu().numpy()}
        if angular:
            fw = [(g[0], g[0], g[2], g[2]) for g in geometries]
            allg = g4 + [f for f in dict.fromkeys(fw) if f not in g4]
            S, Z, xs = tb.scatter(allg, alpha, beta, weight=weight, scale=scale, want_xs=True)
            S, Z, xs = S.cpu().numpy(), Z.cpu().numpy(), xs.cpu().numpy()
            G = len(g4)
            res["S"], res["Z"] = S[:, :G], Z[:, :G]
            res["sca_xsect_h"], res["sca_xsect_v"] = xs[:, :G, 0], xs[:, :G, 1]
            # ext_xsect: optical theorem on the orientation-averaged forward amplitude
            lam = np.broadcast_to(np.asarray(wavelength, dtype=np.float64), (len(radii),))[:, None]
            fi = [allg.index(f) for f in fw]
            res["ext_xsect_h"] = 2.0 * lam * S[:, fi, 1, 1].imag
            res["ext_xsect_v"] = 2.0 * lam * S[:, fi, 0, 0].imag
            inc = list(dict.fromkeys((g[0], g[2]) for g in geometries))
            a = self._asym(tb, inc, alpha, beta, weight, scale, int(res["nmax"].max()))
            gi = [inc.index((g[0], g[2])) for g in geometries]
            res["asym_h"], res["asym_v"] = a[:, gi, 0], a[:, gi, 1]
        else:
            S, Z, _ = tb.scatter(g4, alpha, beta, weight=weight, scale=scale, want_xs=False)
            res["S"], res["Z"] = S.cpu().numpy(), Z.cpu().numpy()
        return res

    # ---- two-layer path (tm2l) ----
    def tm2l(self, d_ext_cm, d_int_cm, ar_ext, ar_int, eps_ext, eps_int, wavelength_cm):
        amp, st = self.eng.tm2l(d_ext_cm, d_int_cm, ar_ext, ar_int, eps_ext, eps_int, wavelength_cm)
        st = st.cpu().numpy()
        if (st != 0).any():
            raise RuntimeError("tm2l: singular matrix for particle %d" % int(np.nonzero(st)[0][0]))
        return amp.cpu().numpy()

    # ---- PSD integration ------------------------------------------------------------------------------------
    def psd_integrate(self, table, D, psd_w, rule="trapz"):
        """table [ncol][n_D], psd_w [n_dsd][n_D] -> [n_dsd][ncol] on the device."""
        wq = trapz_weights(D) if rule == "trapz" else rect_weights(D)
        W = np.asarray(psd_w, dtype=np.float64) * wq[None, :]
        Tt = np.ascontiguousarray(np.asarray(table, dtype=np.float64).T)
        return self.eng.psd_integrate(torch.from_numpy(np.ascontiguousarray(W)), torch.from_numpy(Tt)).cpu().numpy()
