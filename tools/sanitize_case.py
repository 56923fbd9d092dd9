"""Small fused-solver runs for compute-sanitizer: python tools/sanitize_case.py  (periodic + bounded + coastline, a few substeps)"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as e; e.load_package()
import torch
from climaseaice_b200.driver import model_from_case
from climaseaice_b200.synthetic import anticyclone_case, coastline_case, periodic_case
for case in (periodic_case(64, Ny=48, substeps=3, aice="mixed"), anticyclone_case(72, substeps=3), coastline_case(Ny=48, substeps=3)):
    m = model_from_case(case, solver_impl="fused")
    m.update_state()
    m.time_step_momentum(case.dt, 3)
    torch.cuda.synchronize()
    print(case.name, "ok", float(m.all_fields()["u"].parent.abs().max()))
    m.close()
