"""CPU oracles -- TEST INFRASTRUCTURE ONLY.

ctypes bindings of the plain-C restatements under ``oracle/``:

* ``sl_oracle.c``   single-layer EBCM (pytmatrix ``calctmat`` / ``calcampl``)
* ``tm2l_oracle.c`` two-layer EBCM (reference ``hydroscatt/f_2l_tmatrix/tmatrix_py.f``)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  Nothing under ``hydroscatt_b200/`` imports it; the product path has no
CPU fallback.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(force=False):
    """Compile the oracle shared objects with gcc (idempotent)."""
    for name in ("sl_oracle", "tm2l_oracle"):
        src = os.path.join(_HERE, name + ".c")
        out = os.path.join(_HERE, "lib" + name + ".so")
        if not os.path.exists(src):
            continue
        if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
            subprocess.check_call(
                ["gcc", "-O2", "-fPIC", "-std=c11", "-fno-fast-math", "-shared", "-o", out, src, "-lm"])
    # rounding-perturbed twin of the two-layer oracle (FMA contraction on): used only to measure how sensitive the
    # ORACLE's own answer is to last-bit rounding, i.e. to flag ill-conditioned particles (hard part 3)
    for name in ("tm2l_oracle", "sl_oracle"):
        src = os.path.join(_HERE, name + ".c")
        out = os.path.join(_HERE, "lib" + name + "_fma.so")
        if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
            try:
                subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=c11", "-mfma", "-ffp-contract=fast", "-shared",
                                       "-o", out, src, "-lm"])
            except subprocess.CalledProcessError:
                pass


_sl = None


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _sl_proto(lib):
    lib.hso_calctmat.restype = C.c_void_p
    lib.hso_calctmat.argtypes = [C.c_double] * 6 + [C.c_int, C.c_double, C.c_int, C.c_int]
    lib.hso_free.argtypes = [C.c_void_p]
    lib.hso_calcampl.argtypes = [C.c_void_p] + [C.c_double] * 7 + [C.POINTER(C.c_double)] * 2
    return lib


def sl_sensitivity(radius, wavelength, m, axis_ratio, geom=(90.0, 90.0, 0.0, 180.0), ddelt=1e-3, ndgs=2,
                   rescale_ddelt=True):
    """Relative change of the single-layer oracle's own S under last-bit rounding perturbation (FMA-contracted twin
    build): flags ill-conditioned particles (extreme axis ratios), cf. SURVEY.md 7 hard part 3."""
    build()
    path = os.path.join(_HERE, "libsl_oracle_fma.so")
    if not os.path.exists(path):
        return 0.0
    m = complex(m)
    out = []
    for lib in (_sl_proto(C.CDLL(os.path.join(_HERE, "libsl_oracle.so"))), _sl_proto(C.CDLL(path))):
        h = lib.hso_calctmat(float(radius), 1.0, float(wavelength), m.real, m.imag, float(axis_ratio), -1,
                             float(ddelt), int(ndgs), 1 if rescale_ddelt else 0)
        S = np.zeros(8)
        Z = np.zeros(16)
        lib.hso_calcampl(h, float(wavelength), *[float(g) for g in geom[:4]], 0.0, 0.0, _dp(S), _dp(Z))
        lib.hso_free(h)
        out.append(S.view(np.complex128).copy())
    return float(np.abs(out[0] - out[1]).max() / np.abs(out[0]).max())


def sl_lib():
    global _sl
    if _sl is None:
        build()
        lib = C.CDLL(os.path.join(_HERE, "libsl_oracle.so"))
        lib.hso_calctmat.restype = C.c_void_p
        lib.hso_calctmat.argtypes = [C.c_double] * 6 + [C.c_int, C.c_double, C.c_int, C.c_int]
        lib.hso_free.argtypes = [C.c_void_p]
        for f in ("hso_nmax", "hso_ngauss", "hso_ntrials", "hso_status"):
            getattr(lib, f).argtypes = [C.c_void_p]
            getattr(lib, f).restype = C.c_int
        lib.hso_tlen.argtypes = [C.c_void_p]
        lib.hso_tlen.restype = C.c_long
        lib.hso_tdata.argtypes = [C.c_void_p]
        lib.hso_tdata.restype = C.POINTER(C.c_double)
        lib.hso_trials.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                   C.POINTER(C.c_double), C.POINTER(C.c_double)]
        lib.hso_calcampl.argtypes = [C.c_void_p] + [C.c_double] * 7 + [C.POINTER(C.c_double)] * 2
        lib.hso_ampl_from_t.argtypes = [C.POINTER(C.c_double), C.c_int] + [C.c_double] * 7 + \
            [C.POINTER(C.c_double)] * 2
        lib.hso_gauss.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        lib.hso_tm_offset.argtypes = [C.c_int, C.c_int]
        lib.hso_tm_offset.restype = C.c_long
        lib.hso_batch.argtypes = [C.c_int] + [C.POINTER(C.c_double)] * 5 + \
            [C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int] + \
            [C.POINTER(C.c_double)] * 3 + [C.POINTER(C.c_int)] * 4 + [C.POINTER(C.c_double)] * 2
        _sl = lib
    return _sl


class TMatrix:
    """One particle's converged T-matrix from the CPU oracle."""

    def __init__(self, radius, wavelength, m, axis_ratio, rat=1.0, shape=-1, ddelt=1e-3, ndgs=2,
                 rescale_ddelt=True):
        lib = sl_lib()
        m = complex(m)
        self._lib = lib
        self.wavelength = float(wavelength)
        self._h = lib.hso_calctmat(float(radius), float(rat), float(wavelength), m.real, m.imag,
                                   float(axis_ratio), int(shape), float(ddelt), int(ndgs),
                                   1 if rescale_ddelt else 0)
        self.nmax = lib.hso_nmax(self._h)
        self.ngauss = lib.hso_ngauss(self._h)
        self.ntrials = lib.hso_ntrials(self._h)
        self.status = lib.hso_status(self._h)

    def trials(self):
        n = self.ntrials
        a = np.zeros(n, np.int32)
        b = np.zeros(n, np.int32)
        c = np.zeros(n)
        d = np.zeros(n)
        self._lib.hso_trials(self._h, _ip(a), _ip(b), _dp(c), _dp(d))
        return a, b, c, d

    def tdata(self):
        n = self._lib.hso_tlen(self._h)
        p = self._lib.hso_tdata(self._h)
        return np.ctypeslib.as_array(p, shape=(2 * n,)).copy().view(np.complex128)

    def ampl(self, thet0, thet, phi0, phi, alpha=0.0, beta=0.0):
        S = np.zeros(8)
        Z = np.zeros(16)
        self._lib.hso_calcampl(self._h, self.wavelength, float(thet0), float(thet), float(phi0),
                               float(phi), float(alpha), float(beta), _dp(S), _dp(Z))
        return S.view(np.complex128).reshape(2, 2), Z.reshape(4, 4)

    def __del__(self):
        try:
            self._lib.hso_free(self._h)
        except Exception:
            pass


def ampl_from_t(t, nmax, wavelength, thet0, thet, phi0, phi, alpha=0.0, beta=0.0):
    """Amplitude/phase matrices from a packed T-matrix array (complex128, engine layout)."""
    t = np.ascontiguousarray(t, dtype=np.complex128).view(np.float64)
    S = np.zeros(8)
    Z = np.zeros(16)
    sl_lib().hso_ampl_from_t(_dp(t), int(nmax), float(wavelength), float(thet0), float(thet),
                             float(phi0), float(phi), float(alpha), float(beta), _dp(S), _dp(Z))
    return S.view(np.complex128).reshape(2, 2), Z.reshape(4, 4)


def gauss(n):
    z = np.zeros(n)
    w = np.zeros(n)
    sl_lib().hso_gauss(n, _dp(z), _dp(w))
    return z, w


def tm_offset(nmax, m):
    return sl_lib().hso_tm_offset(nmax, m)


def sl_batch(rev, lam, mr, mi, eps, geoms=None, alpha=None, beta=None, wgt=None, shape=-1,
             ddelt=1e-3, ndgs=2, rescale_ddelt=True):
    """Serial CPU batch: returns dict(nmax, ngauss, ntrials, status[, S, Z])."""
    lib = sl_lib()
    rev, lam, mr, mi, eps = [np.ascontiguousarray(np.asarray(a, dtype=np.float64))
                             for a in (rev, lam, mr, mi, eps)]
    n = rev.size
    nmax = np.zeros(n, np.int32)
    ngauss = np.zeros(n, np.int32)
    ntr = np.zeros(n, np.int32)
    st = np.zeros(n, np.int32)
    out = dict(nmax=nmax, ngauss=ngauss, ntrials=ntr, status=st)
    if geoms is not None:
        geoms = np.ascontiguousarray(np.asarray(geoms, dtype=np.float64).reshape(-1, 4))
        alpha = np.ascontiguousarray(np.asarray(alpha, dtype=np.float64))
        beta = np.ascontiguousarray(np.asarray(beta, dtype=np.float64))
        wgt = np.ascontiguousarray(np.asarray(wgt, dtype=np.float64))
        ng = geoms.shape[0]
        S = np.zeros((n, ng, 8))
        Z = np.zeros((n, ng, 16))
        lib.hso_batch(n, _dp(rev), _dp(lam), _dp(mr), _dp(mi), _dp(eps), shape, ddelt, ndgs,
                      1 if rescale_ddelt else 0, ng, _dp(geoms), alpha.size, _dp(alpha), _dp(beta),
                      _dp(wgt), _ip(nmax), _ip(ngauss), _ip(ntr), _ip(st), _dp(S), _dp(Z))
        out["S"] = S.view(np.complex128).reshape(n, ng, 2, 2)
        out["Z"] = Z.reshape(n, ng, 4, 4)
    else:
        null = C.POINTER(C.c_double)()
        lib.hso_batch(n, _dp(rev), _dp(lam), _dp(mr), _dp(mi), _dp(eps), shape, ddelt, ndgs,
                      1 if rescale_ddelt else 0, 0, null, 0, null, null, null, _ip(nmax), _ip(ngauss),
                      _ip(ntr), _ip(st), null, null)
    return out


_t2 = None


def tm2l_lib():
    global _t2
    if _t2 is None:
        build()
        lib = C.CDLL(os.path.join(_HERE, "libtm2l_oracle.so"))
        lib.tm2l_batch.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double)]
        _t2 = lib
    return _t2


def tm2l_sensitivity(d_ext_cm, d_int_cm, ar_ext, ar_int, eps_ext, eps_int, wavelength_cm):
    """Relative change of the oracle's own amplitudes under a last-bit rounding perturbation (FMA-contracted twin
    build).  Particles where this exceeds ~1e-7 are ill-conditioned: no two correct implementations agree to 1e-6."""
    ref = tm2l(d_ext_cm, d_int_cm, ar_ext, ar_int, eps_ext, eps_int, wavelength_cm)
    path = os.path.join(_HERE, "libtm2l_oracle_fma.so")
    if not os.path.exists(path):
        return np.zeros(ref.shape[0])
    twin = tm2l(d_ext_cm, d_int_cm, ar_ext, ar_int, eps_ext, eps_int, wavelength_cm, _lib=C.CDLL(path))
    c = twin[:, 0::2] + 1j * twin[:, 1::2]
    r = ref[:, 0::2] + 1j * ref[:, 1::2]
    return np.abs(c - r).max(axis=1) / np.abs(r).max(axis=1)


def tm2l(d_ext_cm, d_int_cm, ar_ext, ar_int, eps_ext, eps_int, wavelength_cm, _lib=None):
    """Two-layer oracle: returns amp[n, 8] = borrp borip borrpe boripe forrp forip forpe foripe (metres)."""
    arrs = np.broadcast_arrays(np.asarray(d_ext_cm, float), np.asarray(d_int_cm, float), np.asarray(ar_ext, float),
                               np.asarray(ar_int, float), np.asarray(eps_ext, complex), np.asarray(eps_int, complex))
    n = arrs[0].size
    p = np.zeros((n, 8))
    p[:, 0], p[:, 1], p[:, 2], p[:, 3] = [a.reshape(-1) for a in arrs[:4]]
    p[:, 4], p[:, 5] = arrs[4].reshape(-1).real, arrs[4].reshape(-1).imag
    p[:, 6], p[:, 7] = arrs[5].reshape(-1).real, arrs[5].reshape(-1).imag
    out = np.zeros((n, 8))
    lib = _lib if _lib is not None else tm2l_lib()
    lib.tm2l_batch.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double)]
    lib.tm2l_batch(n, _dp(p), float(wavelength_cm), _dp(out))
    return out
