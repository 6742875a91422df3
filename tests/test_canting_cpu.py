"""CPU suite: the numpy restatement of the canting path (tests/canting_ref.py, the checker of the CUDA kernels on the GPU box)
equals the UNMODIFIED reference functions and reproduces the golden fixture written by the unmodified scattering_rain.py."""
import json
import os
import sys

import numpy as np
import pytest

from tests import canting_ref as cr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/hydroscatt"
FX = json.load(open(os.path.join(ROOT, "tests", "golden", "rain_C10.json")))


def _gamma_psd(D, d0, nw, mu, dmax):
    from scipy.special import gamma
    nf = nw * 6.0 / 3.67**4 * (3.67 + mu)**(mu + 4) / gamma(mu + 4)
    d = D[None, :] / d0[:, None]
    v = nf[:, None] * np.exp(mu[:, None] * np.log(d) - (3.67 + mu[:, None]) * d)
    v[:, D > dmax] = 0.0
    return v


def test_restatement_reproduces_golden_fixture():
    D = np.array(FX["D"])
    f = cr.golden_rain_amplitudes(FX)
    a = cr.moments_analytical(10.0)
    sp = cr.canting_sp(53.5, *f, a)
    gold = np.array(FX["sp"])[:, :15]
    assert np.allclose(sp[FX["sp_rows"]], gold, rtol=1e-9, atol=1e-12, equal_nan=True)
    d0, nw, mu = cr.rain_dsd_grid()
    assert d0.size == 9300
    rows = np.array(FX["psd_rows"])
    psd = _gamma_psd(D, d0[rows], nw[rows], mu[rows], 7.0)
    obs = cr.canting_psd(53.5, *f, a, D - np.append(0.0, D[:-1]), psd)
    gold = np.array(FX["psd"])
    cols = FX["psd_columns"][5:]
    for k, name in enumerate(cols):
        assert np.allclose(obs[:, cr.COLS.index(name)], gold[:, 5 + k], rtol=1e-9, atol=1e-12, equal_nan=True), name
    assert np.allclose(np.stack([d0[rows], nw[rows], mu[rows]], 1), gold[:, :3])
    # regression anchor of SURVEY.md section 4: Marshall-Palmer-like DSD D0 = 1.5 mm, Nw = 8000, mu = 0 at C band ~ 40.4 dBZ
    mp = cr.canting_psd(53.5, *f, a, D - np.append(0.0, D[:-1]), _gamma_psd(D, np.array([1.5]), np.array([8000.0]), np.array([0.0]), 7.0))
    assert abs(mp[0, 6] - 40.4) < 0.1, mp[0, 6]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_restatement_equals_unmodified_reference_functions():
    sys.path.insert(0, os.path.join(ROOT, "tests", "stubs"))
    import hydroscatt_b200
    hydroscatt_b200.install_shims()
    sys.path.insert(0, REF)
    try:
        import scattering as ref
    finally:
        sys.path.remove(REF)
    rng = np.random.default_rng(5)
    n = 37
    f = [rng.normal(size=n) * 1e-3 + 1j * rng.normal(size=n) * 1e-4 for _ in range(4)]
    f[3] = f[3].real + 1j * np.abs(f[3].imag) + 1e-5j          # positive extinction
    f[2] = f[2].real + 1j * np.abs(f[2].imag) * 0.5
    for sig in (np.array(10.0), rng.uniform(1.0, 60.0, n)):      # scalar and per-diameter canting
        a_ref = ref.compute_angular_moments_analytical(sig)
        a = cr.moments_analytical(sig)
        for k in a:
            assert np.allclose(a[k], a_ref[k], rtol=1e-15)
        df = ref.compute_scattering_canting_sp(33.3, *f, a_ref)
        assert list(df.columns) == list(cr.COLS)
        assert np.allclose(cr.canting_sp(33.3, *f, a), df.values, rtol=1e-12, atol=0, equal_nan=True)
        dd = rng.uniform(0.05, 0.2, n)
        psd = rng.uniform(0.0, 100.0, (5, n))
        mine = cr.canting_psd(33.3, *f, a, dd, psd)
        for d in range(5):
            r = ref.compute_scattering_canting_psd(33.3, *f, a_ref, dd, psd[d])
            assert list(r.keys()) == list(cr.COLS)
            assert np.allclose(mine[d], [r[k] for k in cr.COLS], rtol=1e-11, atol=0, equal_nan=True)
    A, B = mine[:2], mine[2:4]
    mix = cr.mixture(A, B)
    for d in range(2):
        r = ref.compute_scattering_mixture(dict(zip(cr.COLS, A[d])), dict(zip(cr.COLS, B[d])), var_list=list(cr.COLS))
        for k, v in r.items():
            assert np.isclose(mix[d, cr.COLS.index(k)], v, rtol=1e-12), k
        assert set(cr.COLS) - set(r) == {"sca_xsect_h", "sca_xsect_v", "ext_xsect_h", "ext_xsect_v"}


def test_host_moments_equal_restatement():
    from hydroscatt_b200.canting import compute_angular_moments_analytical, CANT_COLUMNS
    assert CANT_COLUMNS == cr.COLS
    s = np.linspace(0.5, 70.0, 13)
    a, b = compute_angular_moments_analytical(s), cr.moments_analytical(s)
    for k in b:
        assert np.allclose(a[k], b[k], rtol=1e-15)


def test_ziggurat_tables_and_stream_restatement_match_numpy():
    """tools/gen_ziggurat.py: the tables in hydroscatt_b200/csrc/hs_ziggurat.h are the ones of the installed numpy, and the pure-Python
    restatement of PCG64 + random_standard_normal that the device code (hs_cantmc.cuh) follows reproduces numpy draw for draw."""
    import importlib.util
    import os
    import re
    import struct
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_ziggurat", os.path.join(root, "tools", "gen_ziggurat.py"))
    gz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gz)
    ki, wi, fi = gz.tables()
    gz.check(ki, wi, fi, n=30000)
    hdr = open(os.path.join(root, "hydroscatt_b200", "csrc", "hs_ziggurat.h")).read()
    arrays = re.findall(r"(ZIG_\w+)\[256\] = \{([^}]*)\}", hdr)
    got = {name: [int(x.strip().rstrip("UL"), 16) for x in body.split(",")] for name, body in arrays}
    assert got["ZIG_KI"] == list(ki)
    assert got["ZIG_WI_BITS"] == list(struct.unpack("<256Q", struct.pack("<256d", *wi)))
    assert got["ZIG_FI_BITS"] == list(struct.unpack("<256Q", struct.pack("<256d", *fi)))
