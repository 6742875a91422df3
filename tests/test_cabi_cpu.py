"""CPU suite: libhydrotm.so builds for sm_100a, loads, and exports every symbol include/hydrotm.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "hydrotm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hs_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from hydroscatt_b200 import build, _lib
    build.build()
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n
        assert n in _lib.SIGNATURES, f"{n} has no ctypes prototype"
    for n in _lib.SIGNATURES:
        assert n in names, f"{n} bound but not declared in include/hydrotm.h"


def test_host_helpers_without_gpu():
    from hydroscatt_b200 import _lib
    lib = _lib.load()
    assert lib.hs_version().startswith(b"hydrotm-b200")
    assert lib.hs_tmat_size(7) == 2 * (49 + 140)
    assert lib.hs_tmat_mode_offset(7, 0) == 0 and lib.hs_tmat_mode_offset(7, 2) == 2 * (49 + 49)
    assert lib.hs_nmax_start(2.0, 6.5) == 6
    assert lib.hs_nmax_start(0.05, 53.5) == 4


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from hydroscatt_b200 import engine
    with pytest.raises(engine.EngineError):
        engine.Engine(0)
    import hydroscatt_b200
    hydroscatt_b200.install_shims()
    from pytmatrix.tmatrix import Scatterer
    from pytmatrix import tmatrix as tmod
    tmod.set_backend(None)
    with pytest.raises(engine.EngineError):
        Scatterer(radius=1.0, wavelength=33.3, m=7.9 + 2.3j, axis_ratio=1.2).get_SZ()


def test_no_product_import_of_oracle():
    """Nothing under hydroscatt_b200/ may import or load the oracle."""
    pkg = os.path.join(ROOT, "hydroscatt_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "sl_oracle" not in txt, f
