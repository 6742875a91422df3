"""Orientation PDFs and orientation-averaging drivers (pytmatrix.orientation).

The fixed-quadrature average runs as ONE fused device call (all n_alpha x n_beta amplitude evaluations and
their weighted sum) instead of the reference's Python loop over calcampl.
"""
import numpy as np
from scipy.integrate import quad, dblquad


def gaussian_pdf(std=10.0, mean=0.0):
    """Gaussian PDF for orientation averaging (beta in degrees), including the sin(beta) Jacobian."""
    norm_const = 1.0

    def pdf(x):
        return norm_const * np.exp(-0.5 * ((x - mean) / std)**2) * np.sin(np.pi / 180.0 * x)

    norm_dev = quad(pdf, 0.0, 180.0)[0]
    # ensure that the integral over the distribution equals 1
    norm_const /= norm_dev
    return pdf


def uniform_pdf():
    """Uniform (random-orientation) PDF of beta."""
    norm_const = 1.0

    def pdf(x):
        return norm_const * np.sin(np.pi / 180.0 * x)

    norm_dev = quad(pdf, 0.0, 180.0)[0]
    norm_const /= norm_dev
    return pdf


def orient_single(tm):
    """S and Z for a single orientation (tm.alpha, tm.beta)."""
    return tm.get_SZ_single()


def orient_averaged_adaptive(tm):
    """Orientation average by adaptive integration (very slow, as upstream; one device call per sample)."""
    S = np.zeros((2, 2), dtype=complex)
    Z = np.zeros((4, 4))

    def Sfunc(beta, alpha, i, j, real):
        (S_ang, Z_ang) = tm.get_SZ_single(alpha=alpha, beta=beta)
        s = S_ang[i, j].real if real else S_ang[i, j].imag
        return s * tm.or_pdf(beta)

    ind = range(2)
    for i in ind:
        for j in ind:
            S.real[i, j] = dblquad(Sfunc, 0.0, 360.0, lambda x: 0.0, lambda x: 180.0, (i, j, True))[0] / 360.0
            S.imag[i, j] = dblquad(Sfunc, 0.0, 360.0, lambda x: 0.0, lambda x: 180.0, (i, j, False))[0] / 360.0

    def Zfunc(beta, alpha, i, j):
        (S_and, Z_ang) = tm.get_SZ_single(alpha=alpha, beta=beta)
        return Z_ang[i, j] * tm.or_pdf(beta)

    ind = range(4)
    for i in ind:
        for j in ind:
            Z[i, j] = dblquad(Zfunc, 0.0, 360.0, lambda x: 0.0, lambda x: 180.0, (i, j))[0] / 360.0
    return (S, Z)


def fixed_nodes(tm):
    """(alpha[O], beta[O], weight[O], scale) of the n_alpha x n_beta fixed quadrature, alpha-major order."""
    ap = np.linspace(0, 360, tm.n_alpha + 1)[:-1]
    aw = 1.0 / tm.n_alpha
    alpha = np.repeat(ap, len(tm.beta_p))
    beta = np.tile(np.asarray(tm.beta_p, dtype=np.float64), len(ap))
    weight = np.tile(np.asarray(tm.beta_w, dtype=np.float64), len(ap))
    sw = tm.beta_w.sum()
    return alpha, beta, weight, aw / sw


def orient_averaged_fixed(tm):
    """Orientation average with fixed quadrature: n_alpha equispaced alpha x n_beta Gauss points of or_pdf."""
    return tm._orient_averaged_fixed()
