"""BASELINE config 5 at full resolution: coupled slab thermodynamics + EVP dynamics + WENO advection on a 1/12-degree
lat-lon Arctic cap (4320 x 336, phi in (60, 88)), HydrostaticSphericalCoriolis, y-slabs across the ranks.
    python tools/profile_config5.py [nsteps]                                   # one GPU
    python -m torch.distributed.run --nproc-per-node N ... tools/profile_config5.py [nsteps]
Prints ms per time_step! (3 RK stages x {tendencies, 150 substeps, h/aice update, thermodynamics}) and cell-updates/s."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as e; e.load_package()
import torch
import torch.distributed as dist
from climaseaice_b200 import nccl_unique_id
from climaseaice_b200.driver import model_from_case
from climaseaice_b200.synthetic import arctic_cap_case, slab_of

nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
solver = sys.argv[2] if len(sys.argv) > 2 else "auto"
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
K = 4
case = arctic_cap_case(4320, 336, H=7, substeps=150, dt=600.0)
if world > 1:
    sl = slab_of(case, rank, world, 2 * K + 3)
    m = model_from_case(sl, solver_impl=solver, partition=(rank, world, K), device=f"cuda:{local}")
    ids = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    m.comm_init(ids[0])
else:
    m = model_from_case(case, solver_impl=solver, device=f"cuda:{local}")
m.time_step(case.dt)  # warm-up (first step also runs update_state!)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
l0 = m.launch_count
ev0.record()
for _ in range(nsteps):
    m.time_step(case.dt)
ev1.record()
torch.cuda.synchronize()
ms = torch.tensor([ev0.elapsed_time(ev1) / nsteps], device=f"cuda:{local}", dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    cells = case.Nx * case.Ny
    print(f"config5 arctic cap {case.Nx}x{case.Ny} on {world} GPU(s), solver={solver}: {ms.item():.2f} ms per time_step! "
          f"({(m.launch_count - l0) // nsteps} launches), {cells * 450 / ms.item() / 1e6:.3f} G cell-updates/s, stats={m.fused_stats()}", flush=True)
if world > 1:
    dist.destroy_process_group()
