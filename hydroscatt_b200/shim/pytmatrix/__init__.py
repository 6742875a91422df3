"""Drop-in replacement for the `pytmatrix` package as used by openradar/hydroscatt.

Same public surface (Scatterer, orientation, psd, radar, scatter, tmatrix_aux, refractive, quadrature;
SURVEY.md rows P1, P9-P14); the Fortran extension behind it (calctmat / calcampl) is replaced by the
hydroscatt_b200 CUDA engine through its C-ABI.  Put ``hydroscatt_b200/shim`` on ``sys.path`` (or call
``hydroscatt_b200.install_shims()``) and the unmodified hydroscatt scripts import this package.
"""
__version__ = "0.3.2+b200"
