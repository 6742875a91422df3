#!/usr/bin/env python
"""Generates tests/golden/water_refractive_index.json by importing the UNMODIFIED reference module
/root/reference/hydroscatt/refractivity.py (refractive_index_water, Ellison 2005; refractivity.py:59-128) in the
build container.  The table feeds workload W2: 6 bands x T = 0..40 deg C.  The reference tree is
not available on the GPU box, hence the committed fixture."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hydroscatt_b200  # noqa: E402

hydroscatt_b200.install_shims()           # reference part_descrip.py imports pytmatrix.tmatrix_aux
sys.path.insert(0, "/root/reference/hydroscatt")
import refractivity  # noqa: E402

WL = dict(S=111.0, C=53.5, X=33.3, Ku=22.0, Ka=8.43, W=3.19)
out = {"source": "reference hydroscatt/refractivity.py refractive_index_water(wavelength_mm, temp_C)",
       "temps_C": list(range(0, 41)), "bands": WL, "m": {}}
for band, wl in WL.items():
    out["m"][band] = []
    for t in range(0, 41):
        m = complex(refractivity.refractive_index_water(wl, float(t)))
        out["m"][band].append([m.real, m.imag])
path = os.path.join(ROOT, "tests", "golden", "water_refractive_index.json")
json.dump(out, open(path, "w"), indent=0)
print("wrote", path, out["m"]["C"][10])
