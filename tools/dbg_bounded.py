import sys; sys.path.insert(0,'/root/repo')
import __graft_entry__ as e; e.load_package()
import torch
from climaseaice_b200.driver import model_from_case
from climaseaice_b200.synthetic import anticyclone_case
case = anticyclone_case(48, substeps=4)
m = model_from_case(case, solver_impl="fused")
m.update_state()
m.time_step_momentum(case.dt, 4)
torch.cuda.synchronize()
print("ok", float(m.velocities["u"].parent.abs().max()))
