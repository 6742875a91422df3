"""Radar observables from the amplitude / phase matrices (pytmatrix.radar)."""
import numpy as np

from .scatter import ext_xsect


def radar_xsect(scatterer, h_pol=True):
    """Radar cross section for the current setup (horizontal polarization if h_pol)."""
    Z = scatterer.get_Z()
    if h_pol:
        return 2 * np.pi * (Z[0, 0] - Z[0, 1] - Z[1, 0] + Z[1, 1])
    else:
        return 2 * np.pi * (Z[0, 0] + Z[0, 1] + Z[1, 0] + Z[1, 1])


def refl(scatterer, h_pol=True):
    """Reflectivity (with number concentration N=1) for the current setup."""
    return scatterer.wavelength**4 / (np.pi**5 * scatterer.Kw_sqr) * radar_xsect(scatterer, h_pol)


def ldr(scatterer, h_pol=True):
    """Linear depolarization ratio."""
    Z = scatterer.get_Z()
    if h_pol:
        return (Z[0, 0] - Z[0, 1] + Z[1, 0] - Z[1, 1]) / (Z[0, 0] - Z[0, 1] - Z[1, 0] + Z[1, 1])
    else:
        return (Z[0, 0] + Z[0, 1] - Z[1, 0] - Z[1, 1]) / (Z[0, 0] + Z[0, 1] + Z[1, 0] + Z[1, 1])


def Zdr(scatterer):
    """Differential reflectivity Z_h / Z_v (linear)."""
    return radar_xsect(scatterer, True) / radar_xsect(scatterer, False)


def delta_hv(scatterer):
    """Backscatter differential phase delta_hv [rad]."""
    Z = scatterer.get_Z()
    return np.arctan2(Z[2, 3] - Z[3, 2], -Z[2, 2] - Z[3, 3])


def rho_hv(scatterer):
    """Copolarized correlation rho_hv."""
    Z = scatterer.get_Z()
    a = (Z[2, 2] + Z[3, 3])**2 + (Z[3, 2] - Z[2, 3])**2
    b = (Z[0, 0] - Z[0, 1] - Z[1, 0] + Z[1, 1])
    c = (Z[0, 0] + Z[0, 1] + Z[1, 0] + Z[1, 1])
    return np.sqrt(a / (b * c))


def Kdp(scatterer):
    """Specific differential phase K_dp [deg/km]; needs a forward-scattering geometry."""
    if (scatterer.thet0 != scatterer.thet) or (scatterer.phi0 != scatterer.phi):
        raise ValueError("A forward scattering geometry is needed to " +
                         "compute the specific differential phase.")
    S = scatterer.get_S()
    return 1e-3 * (180.0 / np.pi) * scatterer.wavelength * (S[1, 1] - S[0, 0]).real


def Ai(scatterer, h_pol=True):
    """Specific attenuation A [dB/km] for the current setup."""
    return 4.343e-3 * ext_xsect(scatterer, h_pol=h_pol)
