"""CPU suite: the UNMODIFIED hydroscatt scripts run on top of the shims (oracle back-end injected) and reproduce the
committed golden fixture; skipped where the reference tree is absent (e.g. on the GPU box)."""
import json
import os
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/hydroscatt"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_scattering_rain_script_runs_unmodified(tmp_path):
    out = str(tmp_path) + "/"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tests", "run_reference_script.py"), "oracle",
                           "scattering_rain.py", "--band", "C", "--temp", "10", "--path", out,
                           "--analytical_cant_angl", "1"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "rain_C10.json")))
    sp = pd.read_csv(out + "sp_rain_C_1000_ele00000_scattering.csv")
    psd = pd.read_csv(out + "psd_rain_C_1000_ele00000_scattering.csv")
    assert list(sp.columns) == fx["sp_columns"] and list(psd.columns) == fx["psd_columns"]
    assert len(sp) == 70 and len(psd) == 9300
    a, b = sp.iloc[fx["sp_rows"]].values, np.array(fx["sp"])
    assert np.allclose(a, b, rtol=1e-9, atol=1e-12, equal_nan=True)
    a, b = psd.iloc[fx["psd_rows"]].values, np.array(fx["psd"])
    assert np.allclose(a, b, rtol=1e-9, atol=1e-12, equal_nan=True)
    # physical sanity of the regression anchors: 1 mm drop ~ 0 dBZ, 4 mm drop Zdr ~ 2.3 dB
    assert abs(sp["refl_h"][9]) < 0.1 and abs(sp["zdr"][39] - 2.26) < 0.02
