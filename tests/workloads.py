"""Seeded synthetic inputs shared by tests and bench (W1..W5)."""
import numpy as np

WL = dict(S=111.0, C=53.5, X=33.3, Ku=22.0, Ka=8.43, W=3.19)
M_W_10C = dict(S=9.019 + 0.887j, C=8.601 + 1.687j, X=7.942 + 2.332j, Ku=7.042 + 2.777j,
               Ka=4.638 + 2.672j, W=3.117 + 1.665j)


def dsr_thurai_2007(D):
    D = np.asarray(D, dtype=np.float64)
    lo = 1.173 - 0.5165 * D + 0.4698 * D**2 - 0.1317 * D**3 - 8.5e-3 * D**4
    hi = 1.065 - 6.25e-2 * D - 3.99e-3 * D**2 + 7.66e-4 * D**3 - 4.095e-5 * D**4
    return np.where(D < 0.7, 1.0, np.where(D < 1.5, lo, hi))


def rain_grid(band="C", m=None):
    """scattering_rain.py:121-125 diameter grid with Thurai (2007) axis ratios."""
    D = np.arange(0.1, 7.0 + 0.1, 0.1)
    lam = np.full_like(D, WL[band])
    m = M_W_10C[band] if m is None else m
    return dict(D=D, radius=D / 2.0, wavelength=lam, mr=np.full_like(D, m.real), mi=np.full_like(D, m.imag),
                axis_ratio=1.0 / dsr_thurai_2007(D))


def ice_grid(seed=3, n=64):
    """W3-like (scattering_ice_crystals.py regimes): plates / columns, maximum dimension 0.05..2 mm, axis ratios 1.5..6
    (oblate) and 1/4..1/1.5 (prolate; more extreme shapes do not converge in FP64 in either implementation), solid ice at X / Ka / W band.  Equal-volume radius from the maximum dimension as
    Scatterer.equal_volume_from_maximum does."""
    rng = np.random.default_rng(seed)
    L = rng.uniform(0.05, 2.0, n)
    ar = np.where(rng.uniform(size=n) < 0.7, rng.uniform(1.5, 6.0, n), 1.0 / rng.uniform(1.5, 4.0, n))
    band = rng.choice(["X", "Ka", "W"], n)
    lam = np.array([WL[b] for b in band])
    req = np.where(ar > 1, (L / 2) / ar ** (1 / 3), (L / 2) / ar ** (2 / 3))
    return dict(radius=req, wavelength=lam, mr=np.full(n, 1.78), mi=np.full(n, 0.0024), axis_ratio=ar)


def hail_grid(seed=4, n=48):
    """W4-like single-layer hail: D 2..50 mm at C / X / Ka band, dry (ice) to wet (high |m|) refractive indices, mild
    oblateness: NMAX up to ~40, exercises the multi-block assembly and the separate solve kernel."""
    rng = np.random.default_rng(seed)
    D = rng.uniform(2.0, 50.0, n)
    band = rng.choice(["C", "X", "Ka"], n)
    lam = np.array([WL[b] for b in band])
    wet = rng.uniform(size=n)
    mr = 1.78 + wet * 4.0
    mi = 0.003 + wet * 1.5
    ar = 1.0 / rng.uniform(0.7, 1.0, n)
    return dict(radius=D / 2.0, wavelength=lam, mr=mr, mi=mi, axis_ratio=ar)
