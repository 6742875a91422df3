"""CPU suite: the two-layer oracle (oracle/tm2l_oracle.c, a restatement of tmatrix_py.f) against a SECOND, structurally
independent formulation (oracle/tm2l_indep.py: Mishchenko-normalised null-field blocks composed by Peterson-Strom, scipy
Bessel functions, Gauss-Legendre quadrature, LAPACK) on non-spherical particles with DISTINCT layers -- the case no other
anchor covers -- and against the analytic Rayleigh limit of a confocal coated spheroid."""
import numpy as np

import oracle
from oracle import tm2l_indep as ti


def distinct_layer_cases(n=12, seed=3):
    """ar != 1, eps_core != eps_shell, sizes 2 .. 30 mm at S / C / X band (tmatrix_py.f:580-581 input columns)."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        D = rng.uniform(0.2, 3.0)
        f = rng.uniform(0.3, 0.9)
        ar = rng.uniform(0.55, 0.95)
        ar2 = min(1.0, ar + rng.uniform(0, 0.2))
        eo = complex(rng.uniform(2, 9), rng.uniform(0.2, 2.5))**2
        ei = complex(rng.uniform(1.3, 2.0), rng.uniform(0.001, 0.1))**2
        lam = float(rng.choice([5.35, 3.33, 10.0]))
        out.append((D, D * f, ar, ar2, eo, ei, lam))
    return out


def _c(a):
    a = np.asarray(a).ravel()
    return a[0::2] + 1j * a[1::2]


def test_oracle_matches_independent_formulation():
    worst = 0.0
    for c in distinct_layer_cases():
        ref = _c(ti.tm2l(*c))
        got = _c(oracle.tm2l(*[[x] for x in c[:6]], c[6]))
        sens = float(np.max(oracle.tm2l_sensitivity(*[[x] for x in c[:6]], c[6])))
        err = np.abs(got - ref).max() / np.abs(ref).max()
        worst = max(worst, err)
        assert err < max(1e-7, 10 * sens), (c, err, sens)
    print("oracle vs independent formulation, worst relative difference:", worst)


def test_quirk_constants_matter_at_1e7():
    """The float32 pi / float32 1/3 of tmatrix_py.f:161, 617-618 change the answer at the 1e-7 level: the restatement follows
    the quirk, not the exact constants (the two formulations agree 100x better with quirks on than the quirk itself)."""
    c = distinct_layer_cases(1, seed=11)[0]
    q = _c(ti.tm2l(*c, quirks=True))
    e = _c(ti.tm2l(*c, quirks=False))
    o = _c(oracle.tm2l(*[[x] for x in c[:6]], c[6]))
    dq = np.abs(q - o).max() / np.abs(o).max()
    de = np.abs(e - o).max() / np.abs(o).max()
    assert dq < 1e-8 < de < 1e-5, (dq, de)


def test_confocal_coated_spheroid_rayleigh_limit():
    """Electrostatic polarizability of a confocal coated oblate spheroid (Bohren & Huffman eq. 5.35): no T-matrix code at all.
    Both oracles must converge to it as x -> 0 (residual ~ x^2) for a mildly oblate particle."""
    lam_cm = 5.35
    k0 = 2 * ti.CPI32 / (lam_cm / 100.0)
    eps1, eps2 = (3 + 0.5j)**2, (1.78 + 0.002j)**2
    prev = None
    for D in (0.02, 0.01, 0.005):                      # cm
        a_out = (D / 200.0) / 0.95**ti.THIRD32
        a_in = 0.8 * a_out
        svv, shh, ar_in = ti.rayleigh_confocal(a_out, 0.95, a_in, eps1, eps2, k0)
        d_int = 200.0 * a_in * ar_in**ti.THIRD32
        o = _c(oracle.tm2l([D], [d_int], [0.95], [ar_in], [eps1], [eps2], lam_cm))
        i = _c(ti.tm2l(D, d_int, 0.95, ar_in, eps1, eps2, lam_cm))
        for a in (o, i):
            err = max(abs(a[2] - svv) / abs(svv), abs(a[3] - shh) / abs(shh),
                      abs(-a[0] - svv) / abs(svv), abs(a[1] - shh) / abs(shh))
            assert err < 2e-4, (D, err)
        if prev is not None:
            assert err < 0.5 * prev                     # second-order convergence in the size parameter
        prev = err


def test_strongly_oblate_confocal_truncation_converges():
    """ar = 0.7 (confocal core ar 0.45): the fixed nrank = 17 of tmatrix_py.f:549 leaves a 7e-4 truncation error against the
    electrostatic limit, and it shrinks as the order grows -- a property of the method, stated in DESIGN.md."""
    lam_cm = 5.35
    k0 = 2 * np.pi / (lam_cm / 100.0)
    eps1, eps2 = (3 + 0.5j)**2, (1.78 + 0.002j)**2
    D = 0.01
    a_out = (D / 200.0) / 0.7**(1.0 / 3.0)
    a_in = 0.8 * a_out
    svv, shh, ar_in = ti.rayleigh_confocal(a_out, 0.7, a_in, eps1, eps2, k0)
    d_int = 200.0 * a_in * ar_in**(1.0 / 3.0)
    errs = []
    for nmax, ng in ((4, 48), (8, 64), (17, 96)):
        a = _c(ti.tm2l(D, d_int, 0.7, ar_in, eps1, eps2, lam_cm, quirks=False, n_max=nmax, m_max=min(6, nmax), n_gauss=ng))
        errs.append(abs(a[2] - svv) / abs(svv))
    assert errs[0] > errs[1] > errs[2] and errs[2] < 1e-3, errs


def test_single_layer_oracle_matches_independent_formulation():
    """oracle/sl_oracle.c (restated Mishchenko / pytmatrix calctmat + calcampl) against the independent blocks of
    oracle/tm2l_indep.py (scipy Bessel functions, Gauss-Legendre with 96+ nodes, LAPACK) at the oracle's own converged NMAX:
    rain drops at C, X and W band, an oblate ice plate and a prolate column.  The amplitudes are converged to ~ddelt, so the two
    must agree far inside 1e-6 once the oracle is run with a tight ddelt."""
    cases = [(0.5, 53.5, 8.601 + 1.687j, 1.0 / 0.9707), (2.0, 53.5, 8.601 + 1.687j, 1.0 / 0.79), (3.5, 33.3, 7.942 + 2.332j, 1.0 / 0.5577),
             (1.5, 3.19, 3.117 + 1.665j, 1.0 / 0.85), (0.4, 8.43, 1.78 + 0.0024j, 3.0), (0.3, 33.3, 1.78 + 0.0024j, 1.0 / 2.5)]
    worst = 0.0
    for (rad, lam, m, eps) in cases:
        tm = oracle.TMatrix(rad, lam, m, eps, ddelt=1e-9)
        Sb, _ = tm.ampl(90.0, 90.0, 0.0, 180.0)
        Sf, _ = tm.ampl(90.0, 90.0, 0.0, 0.0)
        ref = np.array(ti.spheroid_amplitudes(rad, lam, m, eps, n_max=tm.nmax + 2, n_gauss=max(96, 4 * tm.ngauss)))
        got = np.array([Sb[0, 0], Sb[1, 1], Sf[0, 0], Sf[1, 1]])
        err = np.abs(got - ref).max() / np.abs(ref).max()
        worst = max(worst, err)
        assert err < 1e-6, ((rad, lam, m, eps), tm.nmax, tm.ngauss, err, got, ref)
    print("single-layer oracle vs independent formulation, worst relative difference:", worst)
