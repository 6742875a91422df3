"""Scatterer: T-matrix scattering from nonspherical particles (pytmatrix.tmatrix).

Same attributes, defaults, caching-by-signature and method names as upstream.  The numerics run on the
hydroscatt_b200 CUDA engine: ``_init_tmatrix`` -> hs_calctmat_batch (n=1), ``get_SZ_single`` ->
hs_calcampl_batch, ``orient_averaged_fixed`` -> one fused orientation-reduced hs_calcampl_batch.
"""
import warnings

import numpy as np

from . import orientation
from . import tmatrix_aux
from .psd import PSDIntegrator  # noqa: F401  (upstream re-exports it here)

_backend = None


def get_backend():
    """The numerical back-end: the CUDA engine (created on first use).  There is no CPU fallback."""
    global _backend
    if _backend is None:
        from hydroscatt_b200.backend import CudaBackend
        _backend = CudaBackend()
    return _backend


def set_backend(b):
    """Install another back-end object (used by the CPU-only host-logic tests)."""
    global _backend
    _backend = b


class Scatterer(object):
    """T-Matrix scattering from nonspherical particles.

    Attributes (all settable as constructor kwargs or afterwards):
        radius, radius_type, wavelength, m, axis_ratio, shape, ddelt, ndgs, alpha, beta, thet0, thet,
        phi0, phi, Kw_sqr, orient, or_pdf, n_alpha, n_beta, psd_integrator, psd
    """

    _attr_list = set(["radius", "radius_type", "wavelength", "m", "axis_ratio", "shape", "np", "ddelt",
                      "ndgs", "alpha", "beta", "thet0", "thet", "phi0", "phi", "Kw_sqr", "orient", "scatter",
                      "or_pdf", "n_alpha", "n_beta", "psd_integrator", "psd"])

    _deprecated_aliases = {"axi": "radius", "lam": "wavelength", "eps": "axis_ratio", "rat": "radius_type",
                           "np": "shape", "scatter": "orient"}

    RADIUS_EQUAL_VOLUME = 1.0
    RADIUS_EQUAL_AREA = 0.0
    RADIUS_MAXIMUM = 2.0

    SHAPE_SPHEROID = -1
    SHAPE_CYLINDER = -2
    SHAPE_CHEBYSHEV = 1

    def __init__(self, **kwargs):
        self.radius = 1.0
        self.radius_type = Scatterer.RADIUS_EQUAL_VOLUME
        self.wavelength = 1.0
        self.m = complex(2, 0)
        self.axis_ratio = 1.0
        self.shape = Scatterer.SHAPE_SPHEROID
        self.ddelt = 1e-3
        self.ndgs = 2
        self.alpha = 0.0
        self.beta = 0.0
        self.thet0 = 90.0
        self.thet = 90.0
        self.phi0 = 0.0
        self.phi = 180.0
        self.Kw_sqr = 0.93
        self.orient = orientation.orient_single
        self.or_pdf = orientation.gaussian_pdf()
        self.n_alpha = 5
        self.n_beta = 10
        self.psd_integrator = None
        self.psd = None

        self.nmax = None
        self.ngauss = None
        self._tm = None
        self._tm_signature = ()
        self._single_signature = ()
        self._orient_sz_signature = ()
        self._psd_sz_signature = ()
        self._orient_signature = ()
        self._psd_signature = ()

        for (k, v) in kwargs.items():
            if k in self._deprecated_aliases:
                warnings.warn("'{}' is deprecated, use '{}'".format(k, self._deprecated_aliases[k]),
                              DeprecationWarning)
                k = self._deprecated_aliases[k]
            if k in self._attr_list:
                self.__dict__[k] = v

    def set_geometry(self, geom):
        """Set (thet0, thet, phi0, phi, alpha, beta) from a 6-tuple."""
        (self.thet0, self.thet, self.phi0, self.phi, self.alpha, self.beta) = geom

    def get_geometry(self):
        """The current geometry 6-tuple."""
        return (self.thet0, self.thet, self.phi0, self.phi, self.alpha, self.beta)

    def equal_volume_from_maximum(self):
        if self.shape == Scatterer.SHAPE_SPHEROID:
            if self.axis_ratio > 1.0:  # oblate
                r_eq = self.radius / self.axis_ratio**(1.0 / 3.0)
            else:  # prolate
                r_eq = self.radius / self.axis_ratio**(2.0 / 3.0)
        elif self.shape == Scatterer.SHAPE_CYLINDER:
            if self.axis_ratio > 1.0:  # oblate
                r_eq = self.radius * (0.75 / self.axis_ratio)**(1.0 / 3.0)
            else:  # prolate
                r_eq = self.radius * (0.75 / self.axis_ratio)**(2.0 / 3.0)
        else:
            raise AttributeError("Unsupported shape for maximum radius.")
        return r_eq

    # ---- signatures -------------------------------------------------------------------------------------
    def _tm_sig(self):
        return (self.radius, self.radius_type, self.wavelength, self.m, self.axis_ratio, self.shape,
                self.ddelt, self.ndgs)

    def _set_orient_signature(self):
        self._orient_signature = (self.orient, self.or_pdf, self.n_alpha, self.n_beta)

    def _set_psd_signature(self):
        self._psd_signature = (self.psd,)

    # ---- T-matrix ---------------------------------------------------------------------------------------
    def _radius_args(self):
        if self.radius_type == Scatterer.RADIUS_MAXIMUM:
            # Maximum radius is not directly supported in the original so we convert it to equal volume radius
            return Scatterer.RADIUS_EQUAL_VOLUME, self.equal_volume_from_maximum()
        return self.radius_type, self.radius

    def _init_tmatrix(self):
        """Compute the T-matrix for the current particle parameters (one batched device call, n=1)."""
        (radius_type, radius) = self._radius_args()
        self._tm = get_backend().calctmat(radius, radius_type, self.wavelength, complex(self.m),
                                          self.axis_ratio, self.shape, self.ddelt, self.ndgs)
        self.nmax = self._tm.nmax_host
        self.ngauss = self._tm.ngauss_host
        self._tm_signature = self._tm_sig()
        # a new T-matrix invalidates every amplitude cached from the old one
        self._single_signature = ()
        self._orient_sz_signature = ()

    def _init_orient(self):
        """Gauss points and weights of the orientation PDF for the fixed-quadrature average."""
        if self.orient == orientation.orient_averaged_fixed:
            from . import quadrature
            (self.beta_p, self.beta_w) = quadrature.get_points_and_weights(self.or_pdf, 0, 180, self.n_beta)
        self._set_orient_signature()

    def _ensure_tm(self):
        tm_outdated = self._tm_signature != self._tm_sig()
        if tm_outdated:
            self._init_tmatrix()
        return tm_outdated

    def _ensure_orient(self):
        orient_outdated = self._orient_signature != (self.orient, self.or_pdf, self.n_alpha, self.n_beta)
        if orient_outdated:
            self._init_orient()
        return orient_outdated

    def _orient_nodes(self):
        """(alpha[], beta[], weight[], scale) describing the current orientation treatment."""
        if self.orient == orientation.orient_averaged_fixed:
            self._ensure_orient()
            return orientation.fixed_nodes(self)
        if self.orient == orientation.orient_single:
            return (np.array([self.alpha], dtype=np.float64), np.array([self.beta], dtype=np.float64),
                    np.ones(1), 1.0)
        return None

    def _orient_averaged_fixed(self):
        (alpha, beta, weight, scale) = orientation.fixed_nodes(self)
        return get_backend().orient_avg(self._tm, (self.thet0, self.thet, self.phi0, self.phi), alpha, beta,
                                        weight, scale)

    # ---- amplitude / phase matrices -----------------------------------------------------------------------
    def get_SZ_single(self, alpha=None, beta=None):
        """S and Z for a single orientation."""
        if alpha is None:
            alpha = self.alpha
        if beta is None:
            beta = self.beta
        tm_outdated = self._ensure_tm()
        sig = (self.thet0, self.thet, self.phi0, self.phi, alpha, beta)
        if tm_outdated or self._single_signature != sig:
            (self._S_single, self._Z_single) = get_backend().calcampl(
                self._tm, self.thet0, self.thet, self.phi0, self.phi, alpha, beta)
            self._single_signature = sig
        return (self._S_single, self._Z_single)

    def get_SZ_orient(self):
        """S and Z using the specified orientation averaging."""
        tm_outdated = self._ensure_tm()
        sig = (self.thet0, self.thet, self.phi0, self.phi, self.alpha, self.beta, self.orient)
        orient_outdated = self._ensure_orient()
        if tm_outdated or orient_outdated or self._orient_sz_signature != sig:
            (self._S_orient, self._Z_orient) = self.orient(self)
            self._orient_sz_signature = sig
        return (self._S_orient, self._Z_orient)

    def get_SZ(self):
        """S and Z using the current parameters (PSD-integrated if psd_integrator is set)."""
        if self.psd_integrator is None:
            (self._S, self._Z) = self.get_SZ_orient()
        else:
            sig = (self.thet0, self.thet, self.phi0, self.phi, self.alpha, self.beta, self.orient)
            psd_outdated = self._psd_signature != (self.psd,)
            if psd_outdated or self._psd_sz_signature != sig:
                (self._S, self._Z) = self.psd_integrator(self.psd, self.get_geometry())
                self._psd_sz_signature = sig
                self._set_psd_signature()
        return (self._S, self._Z)

    def get_S(self):
        """The amplitude (S) matrix."""
        return self.get_SZ()[0]

    def get_Z(self):
        """The phase (Z) matrix."""
        return self.get_SZ()[1]

    # ---- device-side replacements of the dblquad-based integrals -------------------------------------------
    def _angular_integrated(self, name, h_pol):
        """sca_xsect / asym for the current incidence direction and orientation treatment."""
        self._ensure_tm()
        nodes = self._orient_nodes()
        if nodes is None:
            raise NotImplementedError("angular integration needs orient_single or orient_averaged_fixed")
        (alpha, beta, weight, scale) = nodes
        res = get_backend().xsect(self._tm, [(self.thet0, self.phi0)], alpha, beta, weight, scale,
                                  want_asym=(name == "asym"))
        key = name + ("_h" if h_pol else "_v")
        return float(res[key][0, 0])

    def _scatter_table(self, radii, m, axis_ratio, geometries, angular):
        """Batched S/Z (+ angular-integrated) table for PSDIntegrator.init_scatter_table."""
        if self.radius_type == Scatterer.RADIUS_MAXIMUM:
            saved = (self.radius, self.axis_ratio)
            r_eq = np.empty_like(radii)
            for i in range(len(radii)):
                (self.radius, self.axis_ratio) = (radii[i], axis_ratio[i])
                r_eq[i] = self.equal_volume_from_maximum()
            (self.radius, self.axis_ratio) = saved
            radius_type, radii = Scatterer.RADIUS_EQUAL_VOLUME, r_eq
        else:
            radius_type = self.radius_type
        nodes = self._orient_nodes()
        if nodes is None:
            raise NotImplementedError("scatter tables need orient_single or orient_averaged_fixed")
        (alpha, beta, weight, scale) = nodes

        def table(geoms, alpha, beta):
            return get_backend().scatter_table(radii, radius_type, self.wavelength, m, axis_ratio, self.shape,
                                               self.ddelt, self.ndgs, geoms, alpha, beta, weight, scale, angular)

        if self.orient != orientation.orient_single:
            return table(geometries, alpha, beta)
        # orient_single: upstream init_scatter_table calls set_geometry(geom) per geometry, so the particle orientation
        # is elements 4, 5 of EACH geometry tuple, not the Scatterer's current alpha / beta
        groups = {}
        for (gi, g) in enumerate(geometries):
            groups.setdefault((float(g[4]), float(g[5])), []).append(gi)
        res = None
        for ((a, b), idx) in groups.items():
            part = table([geometries[i] for i in idx], np.array([a]), np.array([b]))
            if len(groups) == 1:
                return part
            if res is None:
                res = {k: (v if v.ndim < 2 else np.empty((v.shape[0], len(geometries)) + v.shape[2:], v.dtype))
                       for (k, v) in part.items()}
            for (k, v) in part.items():
                if v.ndim >= 2:
                    res[k][:, idx] = v
        return res


# upstream alias
TMatrix = Scatterer
