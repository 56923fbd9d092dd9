"""How many tile passes fall back to the IEEE operators per substep on a mesh with a fold (diagnostic; with a library built with
-DCSI_DEBUG_REDO and CSI_B200_LIB pointing at it, the kernel also prints which tiles)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as e; e.load_package()
import torch
from climaseaice_b200.driver import model_from_case
from climaseaice_b200.synthetic import folded_case
nsub = 2
case = folded_case(96, 80, H=7, substeps=nsub, mask=False)
m = model_from_case(case, solver_impl="fused")
m.update_state()
m.time_step_momentum(case.dt, nsub)
torch.cuda.synchronize()
print("stats (invalid, redone, tiles)", m.fused_stats())
m.close()
