"""Regenerates tests/golden/c_abi_periodic_40x24.bin, the fixture of tests/c_abi_smoke.c (python tests/golden/make_c_abi_fixture.py).

A plain-C caller cannot read .npz, so this fixture is raw little-endian data:
    int32  Nx, Ny, H, nsub;  double dt;
    11 input parents  (u v h a s11 s22 s12 top_x top_y ue ve), each (Ny + 2H) x (Nx + 2H) doubles, i fastest, AFTER update_state!
    5 expected parents (u v s11 s22 s12) after one time_step_momentum!(dt) of nsub substeps under ForwardEuler
The expected arrays come from the CPU oracle (the same seeded periodic case the other fixtures use)."""
import struct
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402

entry.load_package()
from climaseaice_b200.synthetic import periodic_case  # noqa: E402
from tests.helpers import oracle_from_case  # noqa: E402

NX, NY, H, NSUB, DT = 40, 24, 7, 10, 120.0
INPUTS = ("u", "v", "h", "a", "s11", "s22", "s12", "top_x", "top_y", "ue", "ve")
OUTPUTS = ("u", "v", "s11", "s22", "s12")


def main():
    case = periodic_case(NX, Ny=NY, H=H, substeps=NSUB, aice="mixed", timestepper="ForwardEuler")
    o = oracle_from_case(case)
    o.update_state()
    blob = struct.pack("<iiiid", NX, NY, H, NSUB, DT)
    for n in INPUTS:
        a = o.arr[n] if n in o.arr else case.fields[n]
        assert a.shape == (NY + 2 * H, NX + 2 * H), (n, a.shape)
        blob += np.ascontiguousarray(a, dtype="<f8").tobytes()
    o.time_step_momentum(DT, NSUB)
    for n in OUTPUTS:
        blob += np.ascontiguousarray(o.arr[n], dtype="<f8").tobytes()
    out = Path(__file__).resolve().parent / f"c_abi_periodic_{NX}x{NY}.bin"
    out.write_bytes(blob)
    print(out, len(blob), "bytes; max|u| =", float(np.abs(o.arr["u"]).max()))


if __name__ == "__main__":
    main()
