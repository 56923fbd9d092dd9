"""Run a few momentum solves on one GPU (for ncu / quick timing): python tools/profile_case.py N nsub [solver] [bounded|coastline|curvilinear|folded|latlon]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as e; e.load_package()
import torch
from climaseaice_b200.driver import model_from_case
from climaseaice_b200.synthetic import anticyclone_case, periodic_case
N = int(sys.argv[1]); nsub = int(sys.argv[2]); solver = sys.argv[3] if len(sys.argv) > 3 else "auto"
kind = sys.argv[4] if len(sys.argv) > 4 else "periodic"
if kind == "bounded":
    case = anticyclone_case(N, substeps=nsub)
elif kind == "coastline":
    from climaseaice_b200.synthetic import coastline_case
    case = coastline_case(Ny=N, substeps=nsub)
elif kind == "curvilinear":
    from climaseaice_b200.synthetic import curvilinear_case
    case = curvilinear_case(N, N, H=7, substeps=nsub, topology=("Periodic", "Bounded"))
elif kind == "folded":
    from climaseaice_b200.synthetic import folded_case
    case = folded_case(N, N // 2, H=7, substeps=nsub)
elif kind == "latlon":
    from climaseaice_b200.synthetic import latlon_case
    case = latlon_case(N, H=7, substeps=nsub, topology=("Periodic", "Bounded"))
else:
    case = periodic_case(N, substeps=nsub, aice="mixed")
m = model_from_case(case, solver_impl=solver)
m.update_state()
m.time_step_momentum(case.dt, nsub)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(); m.time_step_momentum(case.dt, nsub); ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1)
cells = case.Nx * case.Ny
print(f"{case.name} {case.Nx}x{case.Ny} {solver}: {ms:.3f} ms for {nsub} substeps -> {ms/nsub:.4f} ms/substep, {cells*nsub/ms/1e6:.3f} G cell-updates/s, fused stats (invalid, tile passes redone, tiles) = {m.fused_stats()}")
