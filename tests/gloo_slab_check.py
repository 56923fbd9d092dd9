"""world_size-2 gloo check of the host-side slab partition logic (no GPU)."""
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
from climaseaice_b200 import lib  # noqa: E402
from climaseaice_b200.synthetic import block_of, periodic_case, slab_of  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    K = 4
    Hy = lib().csi_host_halo_width(K)
    assert Hy == 2 * K + 3
    case = periodic_case(24, Ny=16 * world, substeps=K)
    sl = slab_of(case, rank, world, Hy)
    ok = True
    for name, arr in sl.fields.items():
        mine = torch.from_numpy(np.ascontiguousarray(arr[Hy:Hy + sl.Ny]))
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        glob = torch.cat(parts, 0).numpy()
        ok &= np.array_equal(glob, case.fields[name][case.Hy:case.Hy + case.Ny])
        # my north halo is the first Hy interior rows of the next rank (periodic), my south halo the last Hy of the previous
        north = torch.from_numpy(np.ascontiguousarray(arr[Hy + sl.Ny:]))
        first = torch.from_numpy(np.ascontiguousarray(arr[Hy:2 * Hy]))
        recv = [torch.empty_like(first) for _ in range(world)]
        dist.all_gather(recv, first)
        ok &= np.array_equal(north.numpy(), recv[(rank + 1) % world].numpy())
    # 2-D partition world x 1 (blocks along x, rank = rx): interiors tile the case, my east halo is the first Hx interior
    # columns of the next rank (the packed strip csi_exchange_halos sends westwards)
    bcase = periodic_case(16 * world, Ny=12, substeps=K)
    bl = block_of(bcase, rank, world, 1, Hy, Hy)
    for name, arr in bl.fields.items():
        mine = torch.from_numpy(np.ascontiguousarray(arr[Hy:Hy + bl.Ny, Hy:Hy + bl.Nx]))
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        ok &= np.array_equal(torch.cat(parts, 1).numpy(), bcase.fields[name][bcase.Hy:bcase.Hy + bcase.Ny, bcase.Hx:bcase.Hx + bcase.Nx])
        east = np.ascontiguousarray(arr[Hy:Hy + bl.Ny, Hy + bl.Nx:])
        first = torch.from_numpy(np.ascontiguousarray(arr[Hy:Hy + bl.Ny, Hy:2 * Hy]))
        recv = [torch.empty_like(first) for _ in range(world)]
        dist.all_gather(recv, first)
        ok &= np.array_equal(east, recv[(rank + 1) % world].numpy())
    flag = torch.tensor([0 if ok else 1])
    dist.all_reduce(flag)
    if rank == 0:
        print("GLOO_SLABS_OK" if flag.item() == 0 else "GLOO_SLABS_BAD", flush=True)
    dist.destroy_process_group()
    sys.exit(int(flag.item() != 0))


if __name__ == "__main__":
    main()
