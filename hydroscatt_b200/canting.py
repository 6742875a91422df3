"""Batched canting-angle observables: the device equivalent of hydroscatt's scattering.py canting path.

Mirrors the reference's functions (same names, argument meaning and column order) but takes whole batches:

  compute_angular_moments_analytical   scattering.py:136-179  (closed forms, host numpy: six scalars per sigma)
  compute_scattering_canting_sp        scattering.py:182-309  -> one row of 15 observables per diameter
  compute_scattering_canting_psd       scattering.py:312-448  -> one row per DSD, ALL DSDs in one kernel launch
                                       (the reference is called once per DSD: 9 300x in scattering_rain.py:295-316,
                                       11 011x in scattering_hail.py:298-372)
  compute_scattering_mixture           scattering.py:590-673  -> rows of two species mixed incoherently

Everything numerical runs in csrc/hs_cant.cuh through the C-ABI (hs_canting_observables, hs_canting_mixture).
"""
import ctypes as C

import numpy as np
import torch

from . import engine as _engine

#: column order of the reference's DataFrame / dict (scattering.py:277-307, 417-446)
CANT_COLUMNS = ("sca_xsect_h", "sca_xsect_v", "ext_xsect_h", "ext_xsect_v", "ldr_h", "ldr_v", "refl_h", "refl_v", "zdr",
                "rho_hv", "delta_hv", "kdp", "A_h", "A_v", "Adp")
MOMENT_KEYS = ("a1", "a2", "a3", "a4", "a5", "a7")


def compute_angular_moments_analytical(canting_angle):
    """Ryzhkov et al. (2011) closed forms in r = exp(-2 sigma^2) (scattering.py:136-179); dict of arrays."""
    s = np.asarray(canting_angle, dtype=np.float64) * np.pi / 180.0
    r = np.exp(-2.0 * s * s)
    p = 3.0 / 8.0 + 0.5 * r + 1.0 / 8.0 * np.power(r, 4.0)
    q = 3.0 / 8.0 - 0.5 * r + 1.0 / 8.0 * np.power(r, 4.0)
    return {"a1": 0.25 * np.power(1.0 + r, 2.0), "a2": 0.25 * (1.0 - np.power(r, 2.0)), "a3": np.power(p, 2.0), "a4": q * p,
            "a5": 1.0 / 8.0 * p * (1.0 - np.power(r, 4.0)), "a7": 0.5 * r * (1 + r)}


def compute_angular_moments(canting_angle, ele=0.0, n_part=500000, seed=0, eng=None):
    """scattering.compute_angular_moments (scattering.py:70-133) on the device, consuming the same numpy random stream
    (default_rng(seed): one uniform alpha vector, then n_part normal betas per canting angle).  Returns the reference's dict of
    numpy arrays a1, a2, a3, a4, a5, a7; agreement with the reference is at rounding level (libm vs CUDA sin / cos, summation
    order), not at Monte-Carlo level."""
    eng = eng if eng is not None else _engine.default_engine()
    sig = np.atleast_1d(np.asarray(canting_angle, dtype=np.float64)).ravel()
    st = np.random.default_rng(seed=seed).bit_generator.state
    if st["bit_generator"] != "PCG64":
        raise RuntimeError("numpy's default generator is not PCG64")
    s, inc = st["state"]["state"], st["state"]["inc"]
    m64 = (1 << 64) - 1
    state = (C.c_ulonglong * 4)(s >> 64, s & m64, inc >> 64, inc & m64)
    d_sig = eng._dev(sig)
    out = torch.empty((sig.size, 6), dtype=torch.float64, device=eng.device)
    slack = 1.03
    while True:
        nbytes = int(eng.lib.hs_canting_mc_scratch_bytes(sig.size, int(n_part), slack))
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=eng.device)
        with torch.cuda.device(eng.device):
            rc = eng.lib.hs_canting_moments_mc(eng._h, sig.size, _ptr(d_sig), float(ele), int(n_part), state, slack, _ptr(scratch),
                                               nbytes, _ptr(out), eng._stream())
        if rc == 2 and slack < 2.0:
            slack *= 1.5
            continue
        if rc != 0:
            raise _engine.EngineError(f"hs_canting_moments_mc rc={rc}: " + eng._err())
        break
    o = out.cpu().numpy()
    return {k: o[:, i].copy() for i, k in enumerate(MOMENT_KEYS)}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _pack_f(eng, fv180, fh180, fv0, fh0):
    """[n_tab][n_D][8] device tensor from four complex arrays of shape [n_D] or [n_tab, n_D]."""
    fs = [np.atleast_2d(np.asarray(x, dtype=np.complex128)) for x in (fv180, fh180, fv0, fh0)]
    n_tab, n_D = fs[0].shape
    f = np.empty((n_tab, n_D, 8))
    for k, x in enumerate(fs):
        if x.shape != (n_tab, n_D):
            raise ValueError("fv180, fh180, fv0, fh0 must have the same shape")
        f[:, :, 2 * k] = x.real
        f[:, :, 2 * k + 1] = x.imag
    return torch.from_numpy(f).to(eng.device), n_tab, n_D


def _pack_moments(eng, ang_moments_dict, n_tab, n_D):
    """moments -> (device tensor, stride per table, stride per diameter).  Accepted shapes per key: scalar / [1]
    (shared), [n_D] (per diameter: the hail and snow scripts' size-dependent canting), [n_tab, n_D]."""
    a = [np.asarray(ang_moments_dict[k], dtype=np.float64) for k in MOMENT_KEYS]
    shp = np.broadcast_shapes(*[x.shape for x in a])
    m = np.stack([np.broadcast_to(x, shp) for x in a], axis=-1)        # [..., 6]
    if m.size == 6:
        return torch.from_numpy(np.ascontiguousarray(m.reshape(6))).to(eng.device), 0, 0
    if m.shape[:-1] == (n_D,):
        return torch.from_numpy(np.ascontiguousarray(m)).to(eng.device), 0, 6
    if m.shape[:-1] == (n_tab, n_D):
        return torch.from_numpy(np.ascontiguousarray(m)).to(eng.device), 6 * n_D, 6
    raise ValueError("angular moments must be scalars, [n_D] or [n_tab, n_D] arrays")


def _call(eng, f, n_tab, n_D, mom, ms_tab, ms_D, wavelength, kw_sqr, W=None, extra=None):
    lam = eng._dev(wavelength, n_tab)
    if lam.numel() != n_tab:
        raise ValueError("one wavelength per table")
    n_dsd = 0 if W is None else W.shape[0]
    n_extra = 0 if extra is None else extra.shape[1]
    out = torch.empty((n_tab, n_D, 15) if W is None else (n_dsd, n_tab, 15), dtype=torch.float64, device=eng.device)
    out_x = torch.empty((n_dsd, n_extra), dtype=torch.float64, device=eng.device) if n_extra else None
    with torch.cuda.device(eng.device):
        rc = eng.lib.hs_canting_observables(eng._h, n_tab, n_D, _ptr(f), _ptr(mom), ms_tab, ms_D, _ptr(lam), float(kw_sqr), n_dsd,
                                            _ptr(W), n_extra, _ptr(extra), _ptr(out), _ptr(out_x), eng._stream())
    if rc != 0:
        raise _engine.EngineError(f"hs_canting_observables rc={rc}: " + eng._err())
    return out, out_x


def compute_scattering_canting_sp(wavelength, fv180, fh180, fv0, fh0, ang_moments_dict, kw_sqr=0.93, eng=None):
    """Single-particle observables, [n_tab, n_D, 15] device tensor (columns CANT_COLUMNS); inputs [n_D] or [n_tab, n_D]."""
    eng = eng or _engine.default_engine()
    f, n_tab, n_D = _pack_f(eng, fv180, fh180, fv0, fh0)
    mom, ms_tab, ms_D = _pack_moments(eng, ang_moments_dict, n_tab, n_D)
    return _call(eng, f, n_tab, n_D, mom, ms_tab, ms_D, wavelength, kw_sqr)[0]


def compute_scattering_canting_psd(wavelength, fv180, fh180, fv0, fh0, ang_moments_dict, delta_d, psd_vals, kw_sqr=0.93,
                                   extra=None, eng=None):
    """PSD observables for ALL size distributions at once.

    psd_vals: [n_dsd, n_D] numpy array or device tensor (N(D) of every DSD); delta_d: [n_D] bin widths (the rectangle rule of
    scattering_rain.py:271).  Returns obs [n_dsd, n_tab, 15] (and sums [n_dsd, n_extra] of extra [n_D, n_extra] columns weighted
    by psd_vals * delta_d when `extra` is given, e.g. D^3 for precip.compute_lwc)."""
    eng = eng or _engine.default_engine()
    f, n_tab, n_D = _pack_f(eng, fv180, fh180, fv0, fh0)
    mom, ms_tab, ms_D = _pack_moments(eng, ang_moments_dict, n_tab, n_D)
    W = torch.as_tensor(psd_vals, dtype=torch.float64).to(eng.device)
    if W.dim() == 1:
        W = W[None, :]
    if W.shape[1] != n_D:
        raise ValueError("psd_vals must be [n_dsd, n_D]")
    W = (W * eng._dev(delta_d, n_D)[None, :]).contiguous()
    ex = None
    if extra is not None:
        ex = torch.as_tensor(extra, dtype=torch.float64).to(eng.device).reshape(n_D, -1).contiguous()
    obs, sums = _call(eng, f, n_tab, n_D, mom, ms_tab, ms_D, wavelength, kw_sqr, W=W, extra=ex)
    return (obs, sums) if extra is not None else obs


def compute_scattering_mixture(obs_1, obs_2, eng=None):
    """Incoherent mixture of two species: rows [..., 15] in CANT_COLUMNS order -> same shape."""
    eng = eng or _engine.default_engine()
    a = torch.as_tensor(obs_1, dtype=torch.float64).to(eng.device).contiguous()
    b = torch.as_tensor(obs_2, dtype=torch.float64).to(eng.device).contiguous()
    if a.shape != b.shape or a.shape[-1] != 15:
        raise ValueError("mixture: two arrays of identical shape [..., 15]")
    out = torch.empty_like(a)
    with torch.cuda.device(eng.device):
        rc = eng.lib.hs_canting_mixture(eng._h, a.numel() // 15, _ptr(a), _ptr(b), _ptr(out), eng._stream())
    if rc != 0:
        raise _engine.EngineError(f"hs_canting_mixture rc={rc}: " + eng._err())
    return out


def as_frame(obs_rows):
    """dict of numpy columns (what the reference's DataFrame / dict holds) from rows [n, 15]."""
    r = obs_rows.detach().cpu().numpy() if torch.is_tensor(obs_rows) else np.asarray(obs_rows)
    return {k: r[..., i] for i, k in enumerate(CANT_COLUMNS)}
