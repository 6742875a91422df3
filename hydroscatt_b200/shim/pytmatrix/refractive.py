"""Refractive-index tables and mixing rules of pytmatrix.refractive."""
import numpy as np

from . import tmatrix_aux

# complex refractive index of water at the preset wavelengths, 0 / 10 / 20 deg C
m_w_0C = {tmatrix_aux.wl_S: complex(9.075, 1.253), tmatrix_aux.wl_C: complex(8.328, 2.229),
          tmatrix_aux.wl_X: complex(7.351, 2.785), tmatrix_aux.wl_Ku: complex(6.265, 2.993),
          tmatrix_aux.wl_Ka: complex(4.040, 2.388), tmatrix_aux.wl_W: complex(2.880, 1.335)}

m_w_10C = {tmatrix_aux.wl_S: complex(9.019, 0.887), tmatrix_aux.wl_C: complex(8.601, 1.687),
           tmatrix_aux.wl_X: complex(7.942, 2.332), tmatrix_aux.wl_Ku: complex(7.042, 2.777),
           tmatrix_aux.wl_Ka: complex(4.638, 2.672), tmatrix_aux.wl_W: complex(3.117, 1.665)}

m_w_20C = {tmatrix_aux.wl_S: complex(8.876, 0.653), tmatrix_aux.wl_C: complex(8.633, 1.289),
           tmatrix_aux.wl_X: complex(8.208, 1.886), tmatrix_aux.wl_Ku: complex(7.537, 2.424),
           tmatrix_aux.wl_Ka: complex(5.206, 2.801), tmatrix_aux.wl_W: complex(3.382, 1.941)}


def mg_refractive(m, mix):
    """Maxwell-Garnett effective refractive index of a mixture.

    m: tuple of complex refractive indices (matrix first), mix: volume fractions (same length).
    More than two components are merged recursively from the last pair.
    """
    if len(m) == 2:
        cF = float(mix[1]) / (mix[0] + mix[1]) * (m[1]**2 - m[0]**2) / (m[1]**2 + 2 * m[0]**2)
        er = m[0]**2 * (1.0 + 2.0 * cF) / (1.0 - cF)
        m = np.sqrt(er)
    else:
        m_last = mg_refractive(m[-2:], mix[-2:])
        mix_last = mix[-2] + mix[-1]
        m = mg_refractive(m[:-2] + (m_last,), mix[:-2] + (mix_last,))
    return m


def bruggeman_refractive(m, mix):
    """Bruggeman effective-medium refractive index of a two-component mixture."""
    f1 = mix[0] / sum(mix)
    f2 = mix[1] / sum(mix)
    e1 = m[0]**2
    e2 = m[1]**2
    a = -2 * (f1 + f2)
    b = (2 * f1 * e1 - f1 * e2 + 2 * f2 * e2 - f2 * e1)
    c = (f1 + f2) * e1 * e2
    e_eff = (-b - np.sqrt(b**2 - 4 * a * c)) / (2 * a)
    return np.sqrt(e_eff)


ice_density = 0.9167
