#!/usr/bin/env python
"""Generates tests/golden/rain_C10.json in the build container: runs the UNMODIFIED reference script
/root/reference/hydroscatt/scattering_rain.py (--band C --temp 10 --analytical_cant_angl 1) on top of the pytmatrix
shim with the CPU oracle as back-end, and records (a) the single-particle amplitudes the script feeds to
compute_scattering_canting_sp (scattering_rain.py:185-193) and (b) selected rows of the two CSV files it writes.
The reference tree does not exist on the GPU box, hence the committed fixture."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
out = tempfile.mkdtemp() + "/"
subprocess.check_call([sys.executable, os.path.join(ROOT, "tests", "run_reference_script.py"), "oracle",
                       "scattering_rain.py", "--band", "C", "--temp", "10", "--path", out,
                       "--analytical_cant_angl", "1"], stdout=subprocess.DEVNULL)
sp = pd.read_csv(out + "sp_rain_C_1000_ele00000_scattering.csv")
psd = pd.read_csv(out + "psd_rain_C_1000_ele00000_scattering.csv")
import hydroscatt_b200  # noqa: E402
hydroscatt_b200.install_shims()
from pytmatrix import tmatrix as tmod, tmatrix_aux, refractive, orientation  # noqa: E402
from pytmatrix.tmatrix import Scatterer  # noqa: E402
from tests.oracle_backend import OracleBackend  # noqa: E402
tmod.set_backend(OracleBackend())
wl = tmatrix_aux.wl_C
tm = Scatterer(wavelength=wl, m=refractive.m_w_10C[wl])
tm.orient = orientation.orient_single
D = np.arange(0.1, 7.0 + 0.1, 0.1)
Sb, Sf = [], []
for d in D:
    tm.radius = d / 2.0
    tm.axis_ratio = 1.0 / tmatrix_aux.dsr_thurai_2007(d)
    tm.set_geometry(tmatrix_aux.geom_horiz_back)
    S = tm.get_S()
    Sb.append([S[0, 0].real, S[0, 0].imag, S[1, 1].real, S[1, 1].imag])
    tm.set_geometry(tmatrix_aux.geom_horiz_forw)
    S = tm.get_S()
    Sf.append([S[0, 0].real, S[0, 0].imag, S[1, 1].real, S[1, 1].imag])
rows_sp = list(range(0, 70, 3))
rows_psd = list(range(0, len(psd), 311))
fx = {"source": "unmodified reference scattering_rain.py on the pytmatrix shim, CPU oracle back-end; "
                "--band C --temp 10 --analytical_cant_angl 1",
      "D": D.tolist(), "S_back_vv_hh": Sb, "S_forw_vv_hh": Sf,
      "sp_columns": list(sp.columns), "sp_rows": rows_sp, "sp": sp.iloc[rows_sp].values.tolist(),
      "psd_columns": list(psd.columns), "psd_rows": rows_psd, "psd": psd.iloc[rows_psd].values.tolist()}
json.dump(fx, open(os.path.join(ROOT, "tests", "golden", "rain_C10.json"), "w"))
print("wrote rain_C10.json", len(rows_sp), len(rows_psd))
