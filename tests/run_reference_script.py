"""Runs an UNMODIFIED hydroscatt script on top of the shims (test helper).  Usage:
   python tests/run_reference_script.py <backend: oracle|cuda> <script.py> [script args...]"""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/hydroscatt"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "stubs"))
import hydroscatt_b200  # noqa: E402

hydroscatt_b200.install_shims()
backend = sys.argv[1]
script = sys.argv[2]
if backend == "oracle":
    from pytmatrix import tmatrix as tmod
    from tests.oracle_backend import OracleBackend
    tmod.set_backend(OracleBackend())
sys.path.insert(0, REF)
sys.argv = [script] + sys.argv[3:]
os.chdir(REF)
runpy.run_path(os.path.join(REF, script), run_name="__main__")
