"""Regenerates tests/golden/*.npz from the CPU oracle (python tests/golden/make_golden.py).

The reference (Julia) cannot run in this image and ships no golden vectors for this path, so these
fixtures pin the ORACLE's own output: small seeded cases, one full time_step!, every prognostic and
stress field stored bit for bit.  They guard against drift of the oracle (compiler, flags, edits) and
give the GPU tests a committed target that does not depend on the oracle being rebuilt on the box."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402

entry.load_package()
from climaseaice_b200.synthetic import anticyclone_case, coastline_case, curvilinear_case, folded_case, latlon_case, periodic_case  # noqa: E402
from tests.helpers import oracle_from_case  # noqa: E402

CASES = {
    "periodic_24x20_rk3_weno7": lambda: periodic_case(24, Ny=20, substeps=12, aice="mixed"),
    "anticyclone_20_rk3_weno7": lambda: anticyclone_case(20, substeps=12),
    "periodic_18x22_fe_weno5": lambda: periodic_case(18, Ny=22, substeps=9, aice="mixed", advection_order=5, timestepper="ForwardEuler"),
    "latlon_48_rk3_weno7": lambda: latlon_case(48, substeps=20, topology=("Periodic", "Bounded")),          # j-dependent metrics
    "curvilinear_72x56_rk3_weno7": lambda: curvilinear_case(72, 56, substeps=20),                            # (i, j)-dependent metrics
    "coastline_96x48_rk3_weno7": lambda: coastline_case(Ny=48, substeps=16),                                 # immersed coast, drag BC, reduced WENO
    "folded_48x40_rk3_weno7": lambda: folded_case(48, 40, substeps=12),                                      # north fold (copy lists), island
}
FIELDS = ("u", "v", "h", "a", "s11", "s22", "s12", "alpha")


def main():
    out = Path(__file__).resolve().parent
    for name, make in CASES.items():
        case = make()
        o = oracle_from_case(case)
        o.time_step(case.dt)
        np.savez_compressed(out / f"{name}.npz", **{f: o.arr[f] for f in FIELDS})
        print(name, {f: float(np.abs(o.arr[f]).max()) for f in ("u", "s11")})


if __name__ == "__main__":
    main()
