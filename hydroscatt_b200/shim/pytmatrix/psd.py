"""Particle size distributions and PSD integration (pytmatrix.psd).

``PSDIntegrator.init_scatter_table`` builds the whole table with three batched device calls (T-matrices of all
diameters, fused orientation-averaged S/Z for all geometries, angular-integrated cross sections) instead of
the reference's per-diameter Python loop; ``get_SZ`` integrates over D on the device.
"""
import pickle
import warnings
from datetime import datetime

import numpy as np
from scipy.special import gamma

from . import tmatrix_aux


class PSD(object):
    def __call__(self, D):
        if np.shape(D) == ():
            return 0.0
        else:
            return np.zeros_like(D)

    def __eq__(self, other):
        return False

    __hash__ = None


class ExponentialPSD(PSD):
    """Exponential PSD N(D) = N0 exp(-Lambda D), zero above D_max (default 11/Lambda)."""

    def __init__(self, N0=1.0, Lambda=1.0, D_max=None):
        self.N0 = float(N0)
        self.Lambda = float(Lambda)
        self.D_max = 11.0 / Lambda if D_max is None else D_max

    def __call__(self, D):
        psd = self.N0 * np.exp(-self.Lambda * D)
        if np.shape(D) == ():
            if D > self.D_max:
                return 0.0
        else:
            psd[D > self.D_max] = 0.0
        return psd

    def __eq__(self, other):
        try:
            return isinstance(other, ExponentialPSD) and (self.N0 == other.N0) and \
                (self.Lambda == other.Lambda) and (self.D_max == other.D_max)
        except AttributeError:
            return False

    __hash__ = None


class UnnormalizedGammaPSD(ExponentialPSD):
    """Gamma PSD N(D) = N0 D^mu exp(-Lambda D), zero above D_max and at D=0."""

    def __init__(self, N0=1.0, Lambda=1.0, mu=0.0, D_max=None):
        super(UnnormalizedGammaPSD, self).__init__(N0=N0, Lambda=Lambda, D_max=D_max)
        self.mu = mu

    def __call__(self, D):
        # For large mu, this is better numerically than multiplying by D**mu
        with np.errstate(divide="ignore", invalid="ignore"):
            psd = self.N0 * np.exp(self.mu * np.log(D) - self.Lambda * D)
        if np.shape(D) == ():
            if (D > self.D_max) or (D == 0):
                return 0.0
        else:
            psd[(D > self.D_max) | (D == 0)] = 0.0
        return psd

    def __eq__(self, other):
        try:
            return super(UnnormalizedGammaPSD, self).__eq__(other) and self.mu == other.mu
        except AttributeError:
            return False

    __hash__ = None


class GammaPSD(PSD):
    """Normalized gamma PSD N(D) = Nw f(mu) (D/D0)^mu exp(-(3.67+mu) D/D0), zero above D_max (default 3 D0)."""

    def __init__(self, D0=1.0, Nw=1.0, mu=0.0, D_max=None):
        self.D0 = float(D0)
        self.mu = float(mu)
        self.D_max = 3.0 * D0 if D_max is None else D_max
        self.Nw = float(Nw)
        self.nf = Nw * 6.0 / 3.67**4 * (3.67 + mu)**(mu + 4) / gamma(mu + 4)

    def __call__(self, D):
        d = (D / self.D0)
        with np.errstate(divide="ignore", invalid="ignore"):
            psd = self.nf * np.exp(self.mu * np.log(d) - (3.67 + self.mu) * d)
        if np.shape(D) == ():
            if (D > self.D_max) or (D == 0):
                return 0.0
        else:
            psd[(D > self.D_max) | (D == 0.0)] = 0.0
        return psd

    def __eq__(self, other):
        try:
            return isinstance(other, GammaPSD) and (self.D0 == other.D0) and (self.Nw == other.Nw) and \
                (self.mu == other.mu) and (self.D_max == other.D_max)
        except AttributeError:
            return False

    __hash__ = None


class BinnedPSD(PSD):
    """Binned gamma PSD: constant bin_psd[i] on (bin_edges[i], bin_edges[i+1]]."""

    def __init__(self, bin_edges, bin_psd):
        if len(bin_edges) != len(bin_psd) + 1:
            raise ValueError("There must be n+1 bin edges for n bins.")
        self.bin_edges = bin_edges
        self.bin_psd = bin_psd

    def psd_for_D(self, D):
        if not (self.bin_edges[0] < D <= self.bin_edges[-1]):
            return 0.0
        # binary search for the right bin
        start = 0
        end = len(self.bin_edges)
        while end - start > 1:
            half = (start + end) // 2
            if self.bin_edges[start] < D <= self.bin_edges[half]:
                end = half
            else:
                start = half
        return self.bin_psd[start]

    def __call__(self, D):
        if np.shape(D) == ():
            return self.psd_for_D(D)
        else:
            return np.array([self.psd_for_D(d) for d in D])

    def __eq__(self, other):
        if other is None:
            return False
        return len(self.bin_edges) == len(other.bin_edges) and \
            (self.bin_edges == other.bin_edges).all() and (self.bin_psd == other.bin_psd).all()

    __hash__ = None


class PSDIntegrator(object):
    """Integrates scattering properties over a particle size distribution.

    Attributes: num_points (1024), m_func, axis_ratio_func, D_max, geometries.
    """

    attrs = set(["num_points", "m_func", "axis_ratio_func", "D_max", "geometries"])

    def __init__(self, **kwargs):
        self.num_points = 1024
        self.m_func = None
        self.axis_ratio_func = None
        self.D_max = None
        self.geometries = (tmatrix_aux.geom_horiz_back,)

        for k in kwargs:
            if k in self.attrs:
                self.__dict__[k] = kwargs[k]

        self._S_table = None
        self._Z_table = None
        self._angular_table = None
        self._previous_psd = None

    def __call__(self, psd, geometry):
        return self.get_SZ(psd, geometry)

    def _backend(self):
        from . import tmatrix
        return tmatrix.get_backend()

    def get_SZ(self, psd, geometry):
        """S and Z integrated over the PSD (trapezoidal rule over the table diameters)."""
        if (self._S_table is None) or (self._Z_table is None):
            raise AttributeError("Initialize or load the scattering table first.")

        if (not isinstance(psd, PSD)) or self._previous_psd != psd:
            self._S_dict = {}
            self._Z_dict = {}
            psd_w = psd(self._psd_D)
            geoms = list(self.geometries)
            cols = []
            for g in geoms:
                s4 = self._S_table[g].reshape(4, -1)
                cols.append(np.concatenate([np.stack([s4.real, s4.imag], axis=1).reshape(8, -1),
                                            self._Z_table[g].reshape(16, -1)]))
            table = np.concatenate(cols)  # [24*G][n_D]
            out = self._backend().psd_integrate(table, self._psd_D, np.asarray(psd_w, dtype=np.float64)[None, :],
                                                "trapz")[0]
            for (i, geom) in enumerate(geoms):
                o = out[24 * i:24 * (i + 1)]
                self._S_dict[geom] = (o[0:8:2] + 1j * o[1:8:2]).reshape(2, 2)
                self._Z_dict[geom] = o[8:24].reshape(4, 4).copy()
            self._previous_psd = psd

        return (self._S_dict[geometry], self._Z_dict[geometry])

    def get_angular_integrated(self, psd, geometry, property_name, h_pol=True):
        if self._angular_table is None:
            raise AttributeError("Initialize or load the table of angular-integrated quantities first.")

        psd_w = np.asarray(psd(self._psd_D), dtype=np.float64)
        sfx = "" if (h_pol or _UPSTREAM_HPOL_QUIRK) else "_v"

        def integ(y):
            return float(self._backend().psd_integrate(np.asarray(y, dtype=np.float64)[None, :], self._psd_D,
                                                       psd_w[None, :], "trapz")[0, 0])

        def sca_xsect(geom):
            return integ(self._angular_table["sca_xsect" + sfx][geom])

        if property_name == "sca_xsect":
            sca_prop = sca_xsect(geometry)
        elif property_name == "ext_xsect":
            sca_prop = integ(self._angular_table["ext_xsect" + sfx][geometry])
        elif property_name == "asym":
            sca_xsect_int = sca_xsect(geometry)
            if sca_xsect_int > 0:
                sca_prop = integ(self._angular_table["asym" + sfx][geometry] *
                                 self._angular_table["sca_xsect" + sfx][geometry])
                sca_prop /= sca_xsect_int
            else:
                sca_prop = 0.0
        else:
            raise KeyError(property_name)
        return sca_prop

    def init_scatter_table(self, tm, angular_integration=False, verbose=False):
        """Build the S/Z (and optionally angular-integrated) lookup tables over num_points diameters."""
        self._psd_D = np.linspace(self.D_max / self.num_points, self.D_max, self.num_points)

        self._S_table = {}
        self._Z_table = {}
        self._previous_psd = None
        self._m_table = np.empty(self.num_points, dtype=complex)
        if angular_integration:
            self._angular_table = {"sca_xsect": {}, "ext_xsect": {}, "asym": {},
                                   "sca_xsect_v": {}, "ext_xsect_v": {}, "asym_v": {}}
        else:
            self._angular_table = None

        D = self._psd_D
        m = np.array([complex(self.m_func(d)) if self.m_func is not None else complex(tm.m) for d in D])
        ar = np.array([float(self.axis_ratio_func(d)) if self.axis_ratio_func is not None
                       else float(tm.axis_ratio) for d in D])
        self._m_table[:] = m
        if verbose:
            print("Computing {n} points up to D={D} in one batch...".format(n=self.num_points, D=self.D_max))
        res = tm._scatter_table(D / 2.0, m, ar, list(self.geometries), angular_integration)
        for (gi, geom) in enumerate(self.geometries):
            self._S_table[geom] = np.ascontiguousarray(np.moveaxis(res["S"][:, gi], 0, -1))
            self._Z_table[geom] = np.ascontiguousarray(np.moveaxis(res["Z"][:, gi], 0, -1))
            if angular_integration:
                for name in ("sca_xsect", "ext_xsect", "asym"):
                    self._angular_table[name][geom] = res[name + "_h"][:, gi].copy()
                    self._angular_table[name + "_v"][geom] = res[name + "_v"][:, gi].copy()

    def save_scatter_table(self, fn, description=""):
        """Save the lookup tables (same pickle layout as upstream)."""
        data = {
            "description": description,
            "time": datetime.now(),
            "psd_scatter": (self.num_points, self.D_max, self._psd_D, self._S_table, self._Z_table,
                            self._angular_table, self._m_table, self.geometries),
            "version": tmatrix_aux.VERSION,
        }
        pickle.dump(data, open(fn, "wb"), pickle.HIGHEST_PROTOCOL)

    def load_scatter_table(self, fn):
        """Load lookup tables saved by save_scatter_table; returns (time, description)."""
        data = pickle.load(open(fn, "rb"))
        if ("version" not in data) or (data["version"] != tmatrix_aux.VERSION):
            warnings.warn("Loading data saved with another version.", Warning)
        (self.num_points, self.D_max, self._psd_D, self._S_table, self._Z_table, self._angular_table,
         self._m_table, self.geometries) = data["psd_scatter"]
        return (data["time"], data["description"])


# Upstream quirk: with an angular table present, scatter.ext_xsect(tm, h_pol=False) returns
# the h-pol table value, so on the PSD path A_v == A_h and Adp == 0 (scattering.py:576-585).  Default True =
# reference behaviour; set pytmatrix.psd._UPSTREAM_HPOL_QUIRK = False for the physically correct v-pol tables.
_UPSTREAM_HPOL_QUIRK = True
