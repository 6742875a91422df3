"""Constants and helper functions of pytmatrix.tmatrix_aux."""
import numpy as np

VERSION = "0.3.2+b200"

# typical wavelengths [mm] of radar bands
wl_S = 111.0
wl_C = 53.5
wl_X = 33.3
wl_Ku = 22.0
wl_Ka = 8.43
wl_W = 3.19

# typical values of K_w^2 at the bands above
K_w_sqr = {wl_S: 0.93, wl_C: 0.93, wl_X: 0.93, wl_Ku: 0.93, wl_Ka: 0.92, wl_W: 0.75}

# preset geometries (thet0, thet, phi0, phi, alpha, beta), degrees
geom_horiz_back = (90.0, 90.0, 0.0, 180.0, 0.0, 0.0)   # horizontal backscatter
geom_horiz_forw = (90.0, 90.0, 0.0, 0.0, 0.0, 0.0)     # horizontal forward scatter
geom_vert_back = (0.0, 180.0, 0.0, 0.0, 0.0, 0.0)      # vertical backscatter
geom_vert_forw = (180.0, 180.0, 0.0, 0.0, 0.0, 0.0)    # vertical forward scatter


def dsr_thurai_2007(D_eq):
    """Drop shape relationship (vertical/horizontal axis ratio) of Thurai et al. (2007), D_eq in mm."""
    if D_eq < 0.7:
        return 1.0
    elif D_eq < 1.5:
        return 1.173 - 0.5165 * D_eq + 0.4698 * D_eq**2 - 0.1317 * D_eq**3 - 8.5e-3 * D_eq**4
    else:
        return 1.065 - 6.25e-2 * D_eq - 3.99e-3 * D_eq**2 + 7.66e-4 * D_eq**3 - 4.095e-5 * D_eq**4


def dsr_pb(D_eq):
    """Pruppacher and Beard (1970) drop shape relationship."""
    return 1.03 - 0.062 * D_eq


def dsr_bc(D_eq):
    """Beard and Chuang (1987) drop shape relationship."""
    return 1.0048 + 5.7e-04 * D_eq - 2.628e-02 * D_eq**2 + 3.682e-03 * D_eq**3 - 1.677e-04 * D_eq**4
