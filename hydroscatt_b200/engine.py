"""Batched host API over the C-ABI (include/hydrotm.h): torch tensors for device memory and streams only.

``Engine.calctmat`` is the batched equivalent of pytmatrix's ``calctmat`` (one call per particle in the
reference: scattering_rain.py:183-191 via Scatterer._init_tmatrix); ``TMatrixBatch.ampl`` of ``calcampl``
plus ``orientation.orient_averaged_fixed``.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

ST_OK, ST_NMAX_LIMIT, ST_NGAUSS_LIMIT, ST_SINGULAR, ST_ARENA_FULL, ST_BAD_INPUT = 0, 1, 2, 3, 4, 5
STATUS_TEXT = {
    ST_OK: "ok",
    ST_NMAX_LIMIT: "convergence not obtained: NMAX reached NPN1=100",
    ST_NGAUSS_LIMIT: "NGAUSS reached NPNG1=500",
    ST_SINGULAR: "singular Q matrix",
    ST_ARENA_FULL: "T-matrix arena full",
    ST_BAD_INPUT: "invalid particle parameters",
}


class EngineError(RuntimeError):
    pass


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def tmat_size(nmax):
    """complex128 entries of a packed T-matrix (tensor or int nmax)."""
    return 2 * (nmax * nmax + nmax * (nmax + 1) * (2 * nmax + 1) // 6)


class Engine:
    """One CUDA context of the T-matrix engine on ``device``."""

    def __init__(self, device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise EngineError("hydroscatt_b200 needs a CUDA device (no CPU fallback)")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.hs_ctx_create(C.byref(h), self.device.index)
        self._h = h
        if rc != 0:
            raise EngineError("hs_ctx_create: " + self._err())

    def _err(self):
        return self.lib.hs_last_error(self._h).decode()

    def close(self):
        if getattr(self, "_h", None):
            self.lib.hs_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self):
        return int(self.lib.hs_launch_count(self._h))

    def _dev(self, x, n=None, dtype=torch.float64):
        t = torch.as_tensor(x, dtype=dtype)
        if t.device != self.device:
            t = t.to(self.device, non_blocking=True)
        t = t.reshape(-1)
        if n is not None and t.numel() == 1 and n != 1:
            t = t.expand(n)
        return t.contiguous()

    def calctmat_scatter(self, radius, wavelength, m_re, m_im, axis_ratio, geoms, alpha=(0.0,), beta=(0.0,), weight=None,
                         scale=1.0, want_xs=True, **kw):
        """calctmat + TMatrixBatch.scatter in one library call (hs_calctmat_scatter_batch): particles whose T-matrix is
        complete are scattered while the others are still converging.  Returns (TMatrixBatch, S, Z, xs)."""
        g = np.ascontiguousarray(np.asarray(geoms, dtype=np.float64).reshape(-1, 4))
        al = self._dev(alpha)
        be = self._dev(beta)
        w = self._dev(weight if weight is not None else np.ones(al.numel()))
        return self.calctmat(radius, wavelength, m_re, m_im, axis_ratio, _scatter=(g, al, be, w, float(scale), want_xs), **kw)

    def calctmat(self, radius, wavelength, m_re, m_im, axis_ratio, rat=1.0, shape=-1, ddelt=1e-3, ndgs=2,
                 rescale_ddelt=True, arena_margin=1.0, _scatter=None):
        """Converged T-matrices of a batch of spheroids.  All array arguments broadcast to [n]."""
        sizes = [np.size(a) if not torch.is_tensor(a) else a.numel()
                 for a in (radius, wavelength, m_re, m_im, axis_ratio)]
        n = max(sizes)
        if any(k not in (1, n) for k in sizes):
            raise ValueError("calctmat: array arguments must have one element or the batch size %d (got %s)" % (n, sizes))
        axi = self._dev(radius, n)
        lam = self._dev(wavelength, n)
        mr = self._dev(m_re, n)
        mi = self._dev(m_im, n)
        eps = self._dev(axis_ratio, n)
        dev = self.device
        # Arena estimate from the starting truncation and the |m| x predictor of the converged order; the size found for
        # a batch of n particles is remembered, so that steady-state calls need no device->host read at all.
        cap = getattr(self, "_arena_hint", {}).get(n)
        if cap is None:
            xev = 2.0 * np.pi * axi / lam
            n0 = torch.clamp((xev + 4.05 * xev.clamp_min(0) ** 0.333333).nan_to_num(0.0).floor().to(torch.int64), 4, 100)
            a3 = eps.clamp_min(1e-30) ** (1.0 / 3.0)
            pred = (torch.sqrt(mr * mr + mi * mi) * xev * torch.maximum(a3, a3 / eps) + 2.0).nan_to_num(0.0).ceil().to(torch.int64)
            est = tmat_size(torch.clamp(torch.maximum(n0 + 3, torch.minimum(pred, n0 + 12)), max=100))
            cap = int(est.sum().item() * arena_margin) + 1024
        stream = torch.cuda.current_stream(dev)
        while True:
            arena = torch.empty(2 * cap, dtype=torch.float64, device=dev)
            top = torch.zeros(1, dtype=torch.int64, device=dev)
            toff = torch.empty(n, dtype=torch.int64, device=dev)
            nmax = torch.empty(n, dtype=torch.int32, device=dev)
            ngauss = torch.empty(n, dtype=torch.int32, device=dev)
            ntr = torch.empty(n, dtype=torch.int32, device=dev)
            status = torch.empty(n, dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                if _scatter is None:
                    rc = self.lib.hs_calctmat_batch(
                        self._h, n, _ptr(axi), _ptr(lam), _ptr(mr), _ptr(mi), _ptr(eps), float(rat), int(shape),
                        float(ddelt), int(ndgs), 1 if rescale_ddelt else 0, _ptr(arena), cap, _ptr(top), _ptr(toff),
                        _ptr(nmax), _ptr(ngauss), _ptr(ntr), _ptr(status), C.c_void_p(stream.cuda_stream))
                else:
                    g, al, be, w, scale, want_xs = _scatter
                    G = g.shape[0]
                    S = torch.empty((n, G, 2, 2), dtype=torch.complex128, device=dev)
                    Z = torch.empty((n, G, 4, 4), dtype=torch.float64, device=dev)
                    xs = torch.empty((n, G, 2), dtype=torch.float64, device=dev) if want_xs else None
                    rc = self.lib.hs_calctmat_scatter_batch(
                        self._h, n, _ptr(axi), _ptr(lam), _ptr(mr), _ptr(mi), _ptr(eps), float(rat), int(shape),
                        float(ddelt), int(ndgs), 1 if rescale_ddelt else 0, _ptr(arena), cap, _ptr(top), _ptr(toff),
                        _ptr(nmax), _ptr(ngauss), _ptr(ntr), _ptr(status), G, g.ctypes.data_as(C.POINTER(C.c_double)),
                        al.numel(), _ptr(al), _ptr(be), _ptr(w), scale, _ptr(S), _ptr(Z), _ptr(xs),
                        C.c_void_p(stream.cuda_stream))
            if rc == 1:  # HS_RC_ARENA_FULL: the arena counter holds what the batch needs
                cap = max(2 * cap, int(top.item()) + 1024)
                continue
            if rc != 0:
                raise EngineError(f"hs_calctmat_batch rc={rc}: " + self._err())
            break
        if not hasattr(self, "_arena_hint"):
            self._arena_hint = {}
        self._arena_hint[n] = cap
        if _scatter is not None:
            return TMatrixBatch(self, n, arena, top, toff, nmax, ngauss, ntr, status, lam), S, Z, xs
        return TMatrixBatch(self, n, arena, top, toff, nmax, ngauss, ntr, status, lam)

    # ---- PSD integration / observables ---------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def psd_weights(self, kind, p0, p1, p2, dmax, D, wq):
        """W[n_dsd, n_D] = wq[i] * N_d(D_i); kind in {'gamma','exponential','ungamma'} (pytmatrix.psd classes)."""
        k = {"gamma": 0, "exponential": 1, "ungamma": 2}[kind]
        p0 = self._dev(p0)
        nd = p0.numel()
        p1 = self._dev(p1, nd)
        p2 = self._dev(p2 if p2 is not None else 0.0, nd)
        dmax = self._dev(dmax, nd)
        D = self._dev(D)
        wq = self._dev(wq)
        W = torch.empty((nd, D.numel()), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.hs_psd_weights(self._h, k, nd, _ptr(p0), _ptr(p1), _ptr(p2), _ptr(dmax), D.numel(),
                                         _ptr(D), _ptr(wq), _ptr(W), self._stream())
        if rc != 0:
            raise EngineError(f"hs_psd_weights rc={rc}: " + self._err())
        return W

    def psd_integrate(self, W, Tt):
        """out[n_dsd, ncol] = W[n_dsd, n_D] @ Tt[n_D, ncol] (FP64, i ascending)."""
        W = torch.as_tensor(W, dtype=torch.float64).to(self.device).contiguous()
        Tt = torch.as_tensor(Tt, dtype=torch.float64).to(self.device).contiguous()
        nd, nD = W.shape
        assert Tt.shape[0] == nD
        ncol = Tt.shape[1]
        out = torch.empty((nd, ncol), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.hs_psd_integrate(self._h, nd, nD, ncol, _ptr(W), _ptr(Tt), _ptr(out), self._stream())
        if rc != 0:
            raise EngineError(f"hs_psd_integrate rc={rc}: " + self._err())
        return out

    def psd_observables(self, W, rows, n_tab, lam_tab, kw_sqr=0.93, hpol_quirk=True, want_sca=False, columns=None):
        """obs[n_dsd, n_tab, 15] straight from the weights W[n_dsd, n_D] and the table records rows[n_tab*n_D, 56]
        (fused PSD integration + radar observables; the integrated S/Z array is never written).  columns: optional list of
        indices into tables.OBS_COLUMNS -> obs[n_dsd, n_tab, len(columns)] holds only those, in that order."""
        W = W.contiguous()
        rows = rows.contiguous()
        nd, nD = W.shape
        assert rows.shape[0] == n_tab * nD and rows.shape[1] == 56
        lam_tab = self._dev(lam_tab, n_tab)
        with torch.cuda.device(self.device):
            if columns is None:
                obs = torch.empty((nd, n_tab, 15), dtype=torch.float64, device=self.device)
                rc = self.lib.hs_psd_observables(self._h, nd, nD, n_tab, _ptr(W), _ptr(rows), _ptr(lam_tab), float(kw_sqr),
                                                 1 if hpol_quirk else 0, 1 if want_sca else 0, _ptr(obs), self._stream())
            else:
                sel = (C.c_int * len(columns))(*[int(c) for c in columns])
                obs = torch.empty((nd, n_tab, len(columns)), dtype=torch.float64, device=self.device)
                rc = self.lib.hs_psd_observables_sel(self._h, nd, nD, n_tab, _ptr(W), _ptr(rows), _ptr(lam_tab), float(kw_sqr),
                                                     1 if hpol_quirk else 0, len(columns), sel, _ptr(obs), self._stream())
        if rc != 0:
            raise EngineError(f"hs_psd_observables rc={rc}: " + self._err())
        return obs

    def pack_rows(self, S, Z, xs, lam, g_back=0, g_forw=1, f_back=None, f_forw=None, out=None):
        """Table records [n, 56] (tables.py layout) from hs_scatter_batch's S [n, G, 2, 2], Z [n, G, 4, 4], xs [n, G, 2]."""
        n, G = S.shape[0], S.shape[1]
        S, Z, xs = S.contiguous(), Z.contiguous(), xs.contiguous()
        lam = self._dev(lam, n)
        rows = out if out is not None else torch.empty((n, 56), dtype=torch.float64, device=self.device)
        assert rows.is_contiguous() and rows.shape == (n, 56)
        with torch.cuda.device(self.device):
            rc = self.lib.hs_pack_rows(self._h, n, G, _ptr(S), _ptr(Z), _ptr(xs), _ptr(lam), int(g_back), int(g_forw),
                                       int(g_forw if f_back is None else f_back), int(g_forw if f_forw is None else f_forw),
                                       _ptr(rows), self._stream())
        if rc != 0:
            raise EngineError(f"hs_pack_rows rc={rc}: " + self._err())
        return rows

    def gather_rows(self, src, index, out=None):
        """out[i] = src[index[i]] for records of doubles (index: int64 device tensor; negative entries leave out[i] untouched)."""
        src = src.contiguous()
        ncol = src.shape[1]
        index = index.contiguous()
        n_out = index.numel()
        if out is None:
            out = torch.empty((n_out, ncol), dtype=torch.float64, device=self.device)
        assert out.is_contiguous() and out.shape == (n_out, ncol) and index.dtype == torch.int64
        with torch.cuda.device(self.device):
            rc = self.lib.hs_gather_rows(self._h, n_out, ncol, _ptr(src), _ptr(index), _ptr(out), self._stream())
        if rc != 0:
            raise EngineError(f"hs_gather_rows rc={rc}: " + self._err())
        return out

    def observables(self, rows, off_back, off_forw, off_ext, off_sca, lam, kw_sqr=0.93):
        """[n,15] radar observables from rows [n, stride] (layout: include/hydrotm.h hs_observables)."""
        rows = rows.contiguous()
        n, stride = rows.shape
        lam = self._dev(lam)
        lam_stride = 0 if lam.numel() == 1 else 1
        out = torch.empty((n, 15), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.hs_observables(self._h, n, stride, _ptr(rows), off_back, off_forw, off_ext, off_sca,
                                         _ptr(lam), lam_stride, float(kw_sqr), _ptr(out), self._stream())
        if rc != 0:
            raise EngineError(f"hs_observables rc={rc}: " + self._err())
        return out

    def tm2l(self, d_ext_cm, d_int_cm, ar_ext, ar_int, eps_ext, eps_int, wavelength_cm):
        """Two-layer amplitudes [n, 8] (borrp borip borrpe boripe forrp forip forpe foripe, metres) + status [n]."""
        eps_ext = np.asarray(eps_ext, dtype=np.complex128) if not torch.is_tensor(eps_ext) else eps_ext
        eps_int = np.asarray(eps_int, dtype=np.complex128) if not torch.is_tensor(eps_int) else eps_int
        n = max(int(np.size(a)) if not torch.is_tensor(a) else a.numel()
                for a in (d_ext_cm, d_int_cm, ar_ext, ar_int, eps_ext, eps_int))
        de, di = self._dev(d_ext_cm, n), self._dev(d_int_cm, n)
        ae, ai = self._dev(ar_ext, n), self._dev(ar_int, n)
        eor, eoi = self._dev(eps_ext.real, n), self._dev(eps_ext.imag, n)
        eir, eii = self._dev(eps_int.real, n), self._dev(eps_int.imag, n)
        amp = torch.empty((n, 8), dtype=torch.float64, device=self.device)
        status = torch.empty(n, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.hs_tm2l_batch(self._h, n, _ptr(de), _ptr(di), _ptr(ae), _ptr(ai), _ptr(eor), _ptr(eoi),
                                        _ptr(eir), _ptr(eii), float(wavelength_cm), _ptr(amp), _ptr(status),
                                        self._stream())
        if rc != 0:
            raise EngineError(f"hs_tm2l_batch rc={rc}: " + self._err())
        return amp, status

    def fp64_peak(self, tensor=False):
        """Measured FP64 peak of this device, TFLOP/s: DFMA loop, or (tensor=True) DMMA m8n8k4 loop."""
        v = C.c_double(0.0)
        with torch.cuda.device(self.device):
            rc = (self.lib.hs_fp64_tensor_peak if tensor else self.lib.hs_fp64_peak)(self._h, C.byref(v), self._stream())
        if rc != 0:
            raise EngineError(f"hs_fp64_peak rc={rc}: " + self._err())
        return v.value


class TMatrixBatch:
    """Device-resident packed T-matrices of ``n`` particles (layout: include/hydrotm.h)."""

    def __init__(self, eng, n, arena, top, toff, nmax, ngauss, ntrials, status, lam):
        self.eng, self.n = eng, n
        self.arena, self.top, self.toff = arena, top, toff
        self.nmax, self.ngauss, self.ntrials, self.status, self.lam = nmax, ngauss, ntrials, status, lam

    def packed(self, i):
        """complex128 numpy copy of particle i's packed T-matrix."""
        off = int(self.toff[i].item())
        sz = tmat_size(int(self.nmax[i].item()))
        return self.arena[2 * off:2 * (off + sz)].cpu().numpy().view(np.complex128)

    def ampl(self, geoms, alpha=(0.0,), beta=(0.0,), weight=None, scale=1.0, reduce=False):
        """S [n,G,(O,)2,2] complex128 and Z [n,G,(O,)4,4] float64 on the device.

        geoms: [G][4] (thet0, thet, phi0, phi) degrees; alpha/beta: [O] Euler angles (degrees);
        reduce=True returns scale * sum_o weight[o] * (S,Z)(o) (orient_averaged_fixed).
        """
        eng = self.eng
        dev = eng.device
        g = eng._dev(np.asarray(geoms, dtype=np.float64).reshape(-1, 4))
        G = g.numel() // 4
        al = eng._dev(alpha)
        be = eng._dev(beta)
        O = al.numel()
        if be.numel() != O:
            raise ValueError("alpha and beta must have the same length")
        w = eng._dev(weight if weight is not None else np.ones(O))
        if w.numel() != O:
            raise ValueError("weight must have one entry per orientation")
        if reduce:
            S = torch.empty((self.n, G, 2, 2), dtype=torch.complex128, device=dev)
            Z = torch.empty((self.n, G, 4, 4), dtype=torch.float64, device=dev)
        else:
            S = torch.empty((self.n, G, O, 2, 2), dtype=torch.complex128, device=dev)
            Z = torch.empty((self.n, G, O, 4, 4), dtype=torch.float64, device=dev)
        stream = torch.cuda.current_stream(dev)
        with torch.cuda.device(dev):
            rc = eng.lib.hs_calcampl_batch(eng._h, self.n, _ptr(self.arena), _ptr(self.toff), _ptr(self.nmax),
                                           _ptr(self.status), _ptr(self.lam), G, _ptr(g), O, _ptr(al), _ptr(be),
                                           _ptr(w), float(scale), 1 if reduce else 0, _ptr(S), _ptr(Z),
                                           C.c_void_p(stream.cuda_stream))
        if rc != 0:
            raise EngineError(f"hs_calcampl_batch rc={rc}: " + eng._err())
        return S, Z

    def scatter(self, geoms, alpha=(0.0,), beta=(0.0,), weight=None, scale=1.0, want_xs=True):
        """Fused orientation-averaged S [n,G,2,2], Z [n,G,4,4] and (sca_xsect_h, sca_xsect_v) [n,G,2] of each
        geometry's incidence direction: one T-matrix sweep per (orientation, incidence direction)."""
        eng = self.eng
        dev = eng.device
        g = np.ascontiguousarray(np.asarray(geoms, dtype=np.float64).reshape(-1, 4))
        G = g.shape[0]
        al = eng._dev(alpha)
        be = eng._dev(beta)
        O = al.numel()
        w = eng._dev(weight if weight is not None else np.ones(O))
        if be.numel() != O or w.numel() != O:
            raise ValueError("alpha, beta and weight must have the same length")
        S = torch.empty((self.n, G, 2, 2), dtype=torch.complex128, device=dev)
        Z = torch.empty((self.n, G, 4, 4), dtype=torch.float64, device=dev)
        xs = torch.empty((self.n, G, 2), dtype=torch.float64, device=dev) if want_xs else None
        with torch.cuda.device(dev):
            rc = eng.lib.hs_scatter_batch(eng._h, self.n, _ptr(self.arena), _ptr(self.toff), _ptr(self.nmax),
                                          _ptr(self.status), _ptr(self.lam), G,
                                          g.ctypes.data_as(C.POINTER(C.c_double)), O, _ptr(al), _ptr(be), _ptr(w),
                                          float(scale), _ptr(S), _ptr(Z), _ptr(xs), eng._stream())
        if rc != 0:
            raise EngineError(f"hs_scatter_batch rc={rc}: " + eng._err())
        return S, Z, xs

    def xsect(self, inc, alpha=(0.0,), beta=(0.0,), weight=None, scale=1.0):
        """[n, n_inc, 2] orientation-averaged (sca_xsect_h, sca_xsect_v) for incidence directions inc [n_inc][2]."""
        eng = self.eng
        g = eng._dev(np.asarray(inc, dtype=np.float64).reshape(-1, 2))
        G = g.numel() // 2
        al = eng._dev(alpha)
        be = eng._dev(beta)
        O = al.numel()
        w = eng._dev(weight if weight is not None else np.ones(O))
        if be.numel() != O or w.numel() != O:
            raise ValueError("alpha, beta and weight must have the same length")
        out = torch.empty((self.n, G, 2), dtype=torch.float64, device=eng.device)
        with torch.cuda.device(eng.device):
            rc = eng.lib.hs_xsect_batch(eng._h, self.n, _ptr(self.arena), _ptr(self.toff), _ptr(self.nmax),
                                        _ptr(self.status), _ptr(self.lam), G, _ptr(g), O, _ptr(al), _ptr(be),
                                        _ptr(w), float(scale), _ptr(out), eng._stream())
        if rc != 0:
            raise EngineError(f"hs_xsect_batch rc={rc}: " + eng._err())
        return out


_default = {}


def default_engine(device=None):
    """Process-wide engine per device (created on first use)."""
    if not torch.cuda.is_available():
        raise EngineError("hydroscatt_b200 needs a CUDA device (no CPU fallback)")
    idx = torch.cuda.current_device() if device is None else int(device)
    if idx not in _default:
        _default[idx] = Engine(idx)
    return _default[idx]
