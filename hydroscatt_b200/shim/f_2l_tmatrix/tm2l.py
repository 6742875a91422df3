"""Drop-in for the f2py module ``f_2l_tmatrix.tm2l`` built from hydroscatt/f_2l_tmatrix/tmatrix_py.f.

``tmatrix(number_of_angles)`` keeps the Fortran program's side-effect protocol in the current working directory
(reference scattering.py:54-65): it reads ``spongyice.inp`` (namelist ``&ingest palamd=<wavelength cm> /``) and
``fort.20`` (first line N, then N lines ``d_ext d_int ar_ext ar_int Re(eps_ext) Im(eps_ext) Re(eps_int) Im(eps_int)``,
tmatrix_py.f:580-581), and writes ``spongyice.out`` (N lines, Fortran ``8(e15.7)``: back / forward amplitudes,
tmatrix_py.f:2515-2520), ``pck_tmatrix2b.out`` (log) and ``fort.8`` (the duplicate last particle the Fortran driver
leaves behind, tmatrix_py.f:166,470,785-789).  All N particles are solved in ONE batched device call.
"""
import re

import numpy as np

_backend = None


def get_backend():
    global _backend
    if _backend is None:
        from hydroscatt_b200.backend import CudaBackend
        _backend = CudaBackend()
    return _backend


def set_backend(b):
    global _backend
    _backend = b


def fortran_e15_7(x):
    """Fortran ``e15.7`` edit descriptor: ``0.ddddddde+xx`` right-aligned in 15 columns."""
    if x == 0.0 or not np.isfinite(x):
        if not np.isfinite(x):
            return "%15s" % ("NaN" if np.isnan(x) else ("Infinity" if x > 0 else "-Infinity"))
        return "  0.0000000E+00"
    s = "%.6E" % abs(x)  # d.ddddddE+xx
    mant, exp = s.split("E")
    e = int(exp) + 1
    digits = mant.replace(".", "")
    out = "0." + digits + "E%+03d" % e
    if x < 0:
        out = "-" + out
    return "%15s" % out


def read_wavelength_cm(path="spongyice.inp"):
    txt = open(path).read()
    m = re.search(r"palamd\s*=\s*([-+0-9.eEdD]+)", txt)
    if not m:
        raise RuntimeError("tm2l: no palamd in " + path)
    return float(m.group(1).replace("d", "e").replace("D", "e"))


def read_particles(path="fort.20"):
    toks = open(path).read().replace(",", " ").split()
    n = int(float(toks[0]))
    vals = np.array([float(t.replace("d", "e").replace("D", "e")) for t in toks[1:1 + 8 * n]])
    if vals.size != 8 * n:
        raise RuntimeError("tm2l: fort.20 announces %d particles but holds %d values" % (n, vals.size))
    return vals.reshape(n, 8)


def tmatrix(number_of_angles=2):
    """Two-layer T-matrix scattering for every particle listed in ./fort.20 (see module docstring)."""
    if int(number_of_angles) != 2:
        raise RuntimeError("tm2l.tmatrix: only number_of_angles=2 (forward/back) is supported, as hydroscatt uses")
    wl_cm = read_wavelength_cm()
    p = read_particles()
    n = p.shape[0]
    amp = get_backend().tm2l(p[:, 0], p[:, 1], p[:, 2], p[:, 3], p[:, 4] + 1j * p[:, 5], p[:, 6] + 1j * p[:, 7],
                             wl_cm)
    with open("spongyice.out", "w") as f:
        for i in range(n):
            print("on particle %d of %d" % (i + 1, n))
            f.write("".join(fortran_e15_7(v) for v in amp[i]) + "\n")
    print(" writing")
    with open("fort.8", "w") as f:
        f.write("".join(fortran_e15_7(v) for v in amp[n - 1]) + "\n")
    with open("pck_tmatrix2b.out", "w") as f:
        f.write(" two-layer T-matrix (hydroscatt_b200 CUDA engine)\n wavelength [cm] = %g\n particles = %d\n"
                % (wl_cm, n))
        f.write(" nrank = 17  nm = 7  nsect = 2  ib = 8  anginc = 90.0  kgauss = 5\n")
