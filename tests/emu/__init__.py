"""Host emulation of the CUDA device functions (TEST INFRASTRUCTURE ONLY; see emu_driver.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "emu_driver.cpp")
        out = os.path.join(_HERE, "libhs_emu.so")
        csrc = os.path.join(_HERE, "..", "..", "hydroscatt_b200", "csrc")
        deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))]
        if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
            subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-Wno-unknown-pragmas",
                                   "-o", out, src])
        _lib = C.CDLL(out)
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def calctmat(axi, lam, mr, mi, eps, rat=1.0, ddelt=1e-3, ndgs=2, rescale=True, ws_cap=600000, gmax=160,
             arena_cap=None):
    axi, lam, mr, mi, eps = [np.ascontiguousarray(np.atleast_1d(np.asarray(a, dtype=np.float64)))
                             for a in (axi, lam, mr, mi, eps)]
    n = axi.size
    if arena_cap is None:
        arena_cap = n * 200000
    arena = np.zeros(2 * arena_cap)
    toff = np.zeros(n, np.int64)
    nmax = np.zeros(n, np.int32)
    ngauss = np.zeros(n, np.int32)
    ntr = np.zeros(n, np.int32)
    st = np.zeros(n, np.int32)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    lib().emu_calctmat(n, _dp(axi), _dp(lam), _dp(mr), _dp(mi), _dp(eps), C.c_double(rat), C.c_double(ddelt),
                       ndgs, 1 if rescale else 0, _dp(arena), C.c_longlong(arena_cap),
                       toff.ctypes.data_as(C.POINTER(C.c_longlong)), ip(nmax), ip(ngauss), ip(ntr), ip(st),
                       ws_cap, gmax)
    return dict(arena=arena.view(np.complex128), toff=toff, nmax=nmax, ngauss=ngauss, ntrials=ntr, status=st)


def ampl(t, nmax, lam, thet0, thet, phi0, phi, alpha=0.0, beta=0.0):
    t = np.ascontiguousarray(t, dtype=np.complex128).view(np.float64)
    out = np.zeros(24)
    lib().emu_ampl(_dp(t), int(nmax), C.c_double(lam), C.c_double(thet0), C.c_double(thet), C.c_double(phi0),
                   C.c_double(phi), C.c_double(alpha), C.c_double(beta), _dp(out))
    return out[:8].view(np.complex128).reshape(2, 2), out[8:].reshape(4, 4)


def xsect(t, nmax, lam, thet0, phi0, alpha=0.0, beta=0.0):
    t = np.ascontiguousarray(t, dtype=np.complex128).view(np.float64)
    out = np.zeros(2)
    lib().emu_xsect(_dp(t), int(nmax), C.c_double(lam), C.c_double(thet0), C.c_double(phi0), C.c_double(alpha),
                    C.c_double(beta), _dp(out))
    return out


def tm2l(p8, wl_cm):
    p8 = np.ascontiguousarray(p8, dtype=np.float64).reshape(-1, 8)
    n = p8.shape[0]
    amp = np.zeros((n, 8))
    st = np.zeros(n, np.int32)
    lib().emu_tm2l(n, _dp(p8), C.c_double(wl_cm), _dp(amp), st.ctypes.data_as(C.POINTER(C.c_int)))
    return amp, st


def scatter(t, nmax, lam, thet0, phi0, gsc, alpha=0.0, beta=0.0):
    """fused device function: returns (S[ng,2,2], Z[ng,4,4], xs[2]) for scattered directions gsc [ng][2]."""
    t = np.ascontiguousarray(t, dtype=np.complex128).view(np.float64)
    gsc = np.ascontiguousarray(gsc, dtype=np.float64).reshape(-1, 2)
    ng = gsc.shape[0]
    out = np.zeros(ng * 24)
    xs = np.zeros(2)
    lib().emu_scatter(_dp(t), int(nmax), C.c_double(lam), C.c_double(thet0), C.c_double(phi0), ng, _dp(gsc),
                      C.c_double(alpha), C.c_double(beta), _dp(out), _dp(xs))
    o = out.reshape(ng, 24)
    return o[:, :8].copy().view(np.complex128).reshape(ng, 2, 2), o[:, 8:].reshape(ng, 4, 4), xs


def gauss_full(n):
    z, w = np.zeros(n), np.zeros(n)
    lib().emu_gauss_full(int(n), _dp(z), _dp(w))
    return z, w


def gauss_positive(n):
    x, w = np.zeros(n // 2), np.zeros(n // 2)
    lib().emu_gauss_positive(int(n), _dp(x), _dp(w))
    return x, w


def dv_rounds(s0, lb, table, ndgs=2, spec=0, spec0=0, margin=1, ddelt=1e-4):
    """Device-driven rounds of ONE particle (controller + planner of hs_dvctl.h as host code) fed from a table of trial results
    table = (nmax[], ngauss[], qsca[], qext[]).  Returns dict(nmax, ngauss, ntrials, status, done, rounds, planned, unknown)."""
    tn, tg, ts, te = [np.ascontiguousarray(a) for a in table]
    tn = tn.astype(np.int32); tg = tg.astype(np.int32); ts = ts.astype(np.float64); te = te.astype(np.float64)
    out = np.zeros(8, np.int32)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    lib().emu_dv_rounds(int(s0), int(lb), int(ndgs), int(spec), int(spec0), int(margin), C.c_double(ddelt), int(tn.size),
                        ip(tn), ip(tg), _dp(ts), _dp(te), ip(out))
    return dict(zip(("nmax", "ngauss", "ntrials", "status", "done", "rounds", "planned", "unknown"), [int(v) for v in out]))
