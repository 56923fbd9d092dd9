import sys, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as e; e.load_package()
from climaseaice_b200 import lib
for span in (0, 5, 30, 250, 600):
    out = (C.c_uint64 * 5)()
    rc = lib().csi_selftest_math(200_000_000, 7 + span, span, out)
    print("span", span, "rc", rc, "rcp/div/sqrt/divc mismatches", list(out)[:4], "rejected", out[4])
