"""Small fused-solver runs for compute-sanitizer: python tools/sanitize_case.py  (periodic, bounded, coastline, lat-lon cap, the
marginal-ice variants -- free drift, top drag, prescribed bottom stress --, a curvilinear mesh and a mesh with a fold, a few substeps
each, plus one full time step with WENO next to an immersed coast)"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as e; e.load_package()
import torch
from climaseaice_b200.driver import model_from_case
from climaseaice_b200.synthetic import anticyclone_case, arctic_cap_case, coastline_case, curvilinear_case, folded_case, marginal_ice_case, periodic_case
cases = [periodic_case(64, Ny=48, substeps=3, aice="mixed"), anticyclone_case(72, substeps=3), coastline_case(Ny=48, substeps=3),
         arctic_cap_case(96, 40, substeps=3)]
cases += [marginal_ice_case(40, substeps=3, variant=v) for v in ("bottom_drag", "top_drag", "fields", "both_drag", "const_top_drag")]
# two-dimensional metric planes; a fold (split substep next to it); regular grids run the cooperative multi-substep launch (3 substeps:
# two of them in it), the others one launch per substep
cases += [curvilinear_case(72, 56, H=5, substeps=3, topology=("Periodic", "Bounded")), folded_case(64, 56, H=5, substeps=3)]
for case in cases:
    m = model_from_case(case, solver_impl="fused")
    m.update_state()
    m.time_step_momentum(case.dt, 3)
    torch.cuda.synchronize()
    print(case.name, "ok", float(m.all_fields()["u"].parent.abs().max()), m.fused_stats())
    m.close()
c = coastline_case(Ny=32, substeps=2)
m = model_from_case(c, solver_impl="fused")
m.time_step(c.dt)
torch.cuda.synchronize()
print("coastline full step ok")
m.close()
