"""Oracle-backed stand-in for hydroscatt_b200.backend.CudaBackend (TEST INFRASTRUCTURE ONLY): lets the CPU test
suite exercise the shim's host logic (caching, orientation nodes, PSD classes, table packing) without a GPU."""
import numpy as np

import oracle
from tests import emu
from tests.layout import packed_size


class _H:
    def __init__(self, tm, lam):
        self.tm, self.lam = tm, lam
        self.nmax_host, self.ngauss_host, self.status_host = tm.nmax, tm.ngauss, tm.status

    @property
    def nmax(self):
        return self.nmax_host


def _trapz_w(D):
    w = np.zeros_like(D)
    d = np.diff(D)
    w[:-1] += 0.5 * d
    w[1:] += 0.5 * d
    return w


class OracleBackend:
    def calctmat(self, radius, radius_type, wavelength, m, axis_ratio, shape, ddelt, ndgs):
        tm = oracle.TMatrix(radius, wavelength, m, axis_ratio, rat=radius_type, shape=shape, ddelt=ddelt, ndgs=ndgs)
        if tm.status not in (0, 2):
            raise RuntimeError("oracle calctmat status %d" % tm.status)
        return _H(tm, wavelength)

    def calcampl(self, h, thet0, thet, phi0, phi, alpha, beta):
        return h.tm.ampl(thet0, thet, phi0, phi, alpha, beta)

    def orient_avg(self, h, geom4, alpha, beta, weight, scale):
        S = np.zeros((2, 2), dtype=complex)
        Z = np.zeros((4, 4))
        for a, b, w in zip(alpha, beta, weight):
            s, z = h.tm.ampl(*geom4, a, b)
            S += w * s
            Z += w * z
        return S * scale, Z * scale

    def xsect(self, h, inc, alpha, beta, weight, scale, want_asym=False):
        raise NotImplementedError("the CPU oracle has no closed-form sca_xsect; the GPU tests check the CUDA one against "
                                  "brute-force quadrature of oracle amplitudes (tests/test_scatter_gpu.py)")

    def psd_integrate(self, table, D, psd_w, rule="trapz"):
        D = np.asarray(D, dtype=np.float64)
        wq = _trapz_w(D) if rule == "trapz" else D - np.concatenate([[0.0], D[:-1]])
        return (np.asarray(psd_w) * wq[None, :]) @ np.asarray(table).T

    def scatter_table(self, radii, radius_type, wavelength, m, axis_ratio, shape, ddelt, ndgs, geometries, alpha,
                      beta, weight, scale, angular):
        n, G = len(radii), len(geometries)
        S = np.zeros((n, G, 2, 2), dtype=complex)
        Z = np.zeros((n, G, 4, 4))
        res = {"S": S, "Z": Z}
        if angular:
            for k in ("sca_xsect_h", "sca_xsect_v", "ext_xsect_h", "ext_xsect_v", "asym_h", "asym_v"):
                res[k] = np.full((n, G), np.nan)
        for i in range(n):
            h = self.calctmat(radii[i], radius_type, wavelength, complex(m[i]), axis_ratio[i], shape, ddelt, ndgs)
            for gi, g in enumerate(geometries):
                S[i, gi], Z[i, gi] = self.orient_avg(h, tuple(g[:4]), alpha, beta, weight, scale)
                if angular:
                    sf, _ = self.orient_avg(h, (g[0], g[0], g[2], g[2]), alpha, beta, weight, scale)
                    res["ext_xsect_h"][i, gi] = 2 * wavelength * sf[1, 1].imag
                    res["ext_xsect_v"][i, gi] = 2 * wavelength * sf[0, 0].imag
        return res
