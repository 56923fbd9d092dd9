"""N-GPU == 1-GPU check (the analogue of test/distributed_tests_utils.jl:40-88), launched by
tests/test_multigpu.py through torch.distributed.run.  Every rank steps its y-slab of a doubly
periodic case (NCCL halo exchange every K substeps) and compares with the same global case stepped
on one GPU.  Exit code 0 = bitwise equal."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402

entry.load_package()
from climaseaice_b200 import nccl_unique_id  # noqa: E402
from climaseaice_b200.driver import model_from_case  # noqa: E402
from climaseaice_b200.synthetic import arctic_cap_case, block_of, periodic_case, slab_of  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    solver = sys.argv[1] if len(sys.argv) > 1 else "auto"
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    topo = sys.argv[4] if len(sys.argv) > 4 else "periodic"
    Rx = int(sys.argv[5]) if len(sys.argv) > 5 else 1      # 2-D partition Rx x (world / Rx), rank = ry * Rx + rx
    Ry = world // Rx
    case = periodic_case(96 * Rx, Ny=32 * Ry, substeps=10, aice="mixed")
    if topo == "bounded_x":   # Bounded x Periodic: walls at the west of column 0 and the east of the last column of ranks
        case.topology = ("Bounded", "Periodic")
        case.v_bc_value = 0.0
        for k in ("u", "top_x", "ue"):
            case.fields[k] = np.ascontiguousarray(np.hstack([case.fields[k], case.fields[k][:, -1:]]))
    if topo == "bounded_y":   # Periodic x Bounded: walls at the south of rank 0 and the north of the last rank
        case.topology = ("Periodic", "Bounded")
        case.u_bc_value = 0.0
        for k in ("v", "top_y", "ve"):
            case.fields[k] = np.ascontiguousarray(np.vstack([case.fields[k], case.fields[k][-1:]]))
    if topo == "coastline":   # BASELINE config 4 in miniature: Periodic x Bounded, immersed coast, linear immersed drag, uniform wind
        from climaseaice_b200.synthetic import coastline_case
        case = coastline_case(Ny=24 * Ry, H=7, substeps=10)
        assert case.Nx % Rx == 0
    if topo == "curvilinear":  # two-dimensional metrics (orthogonal curvilinear mesh), doubly periodic, general kernels
        from climaseaice_b200.synthetic import curvilinear_case
        case = curvilinear_case(64, 32 * Ry, H=7, substeps=10, topology=("Periodic", "Periodic"))
    if topo == "folded":       # a tripolar-like mesh: two-dimensional metrics, a fold in the north (held by the last slab), an island at the fold
        from climaseaice_b200.synthetic import folded_case
        case = folded_case(64, 32 * Ry, H=7, substeps=10)
    if topo == "arctic":      # BASELINE config 5 in miniature: lat-lon cap, coupled thermodynamics, zonally periodic, walls in y
        case = arctic_cap_case(96, 24 * world, H=7, substeps=10)
    Hy = max(2 * K + 3, 7)
    if Rx > 1:
        sl = block_of(case, rank, Rx, Ry, Hy, Hy)
        m = model_from_case(sl, solver_impl=solver, partition=(rank, world, K, Rx), device=f"cuda:{local}")
    else:
        sl = slab_of(case, rank, world, Hy)
        m = model_from_case(sl, solver_impl=solver, partition=(rank, world, K), device=f"cuda:{local}")
    ids = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    m.comm_init(ids[0])
    ref = model_from_case(case, solver_impl=solver, device=f"cuda:{local}")
    for _ in range(nsteps):
        m.time_step(case.dt)
        ref.time_step(case.dt)
    torch.cuda.synchronize()
    ny, nx = sl.Ny, sl.Nx
    rx, ry = rank % Rx, rank // Rx
    ok = True
    for n in ("u", "v", "h", "a", "s11", "s22", "s12"):
        mine = m.all_fields()[n].parent[Hy:Hy + ny, sl.Hx:sl.Hx + sl.Nx]
        glob = ref.all_fields()[n].parent[case.Hy + ry * ny:case.Hy + (ry + 1) * ny, case.Hx + rx * nx:case.Hx + (rx + 1) * nx]
        if not torch.equal(mine, glob):
            ok = False
            d = (mine - glob).abs().max().item()
            print(f"rank {rank}: field {n} differs, max abs {d:.3e}", flush=True)
    flag = torch.tensor([0 if ok else 1], device=f"cuda:{local}")
    dist.all_reduce(flag)
    if rank == 0:
        print("MULTIGPU_OK" if flag.item() == 0 else "MULTIGPU_MISMATCH", f"world={world} solver={solver} K={K} topo={topo} Rx={Rx}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
