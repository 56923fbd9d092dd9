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
from climaseaice_b200.synthetic import block_of, coastline_case, curvilinear_case, folded_case, periodic_case, slab_of  # noqa: E402


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
    # immersed mask of a partitioned case (BASELINE config 4 in miniature): the slabs' interior rows tile the global mask, and a
    # slab's halo rows are its neighbours' rows (Bounded y: nothing immersed beyond the global parent)
    cc = coastline_case(Ny=16 * world, H=7, substeps=K)
    cs = slab_of(cc, rank, world, Hy)
    mine = torch.from_numpy(np.ascontiguousarray(cs.mask[Hy:Hy + cs.Ny]))
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    ok &= np.array_equal(torch.cat(parts, 0).numpy(), cc.mask[cc.Hy:cc.Hy + cc.Ny])
    if rank + 1 < world:
        ok &= np.array_equal(cs.mask[Hy + cs.Ny:], parts[rank + 1].numpy()[:Hy])
    cb = block_of(cc, rank, world, 1, Hy, Hy)          # the same case cut along its periodic x axis
    mine = torch.from_numpy(np.ascontiguousarray(cb.mask[Hy:Hy + cb.Ny, Hy:Hy + cb.Nx]))
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    ok &= np.array_equal(torch.cat(parts, 1).numpy(), cc.mask[cc.Hy:cc.Hy + cc.Ny, cc.Hx:cc.Hx + cc.Nx])
    ok &= np.array_equal(cb.mask[Hy:Hy + cb.Ny, Hy + cb.Nx:], parts[(rank + 1) % world].numpy()[:, :Hy])
    # two-dimensional metrics on y-slabs: local row j of a slab is global row rank * ny + j of the (exactly periodic) global arrays
    mc = curvilinear_case(24, 16 * world, H=7, substeps=K, topology=("Periodic", "Periodic"))
    ms = slab_of(mc, rank, world, Hy)
    for name, G in mc.metric_arrays.items():
        loc = ms.metric_arrays[name]
        ok &= loc.shape == (ms.Ny + 2 * Hy + 1, mc.Nx + 2 * mc.Hx + 1)
        for jl in range(1 - Hy, ms.Ny + Hy + 2):
            jg = (rank * ms.Ny + jl - 1) % mc.Ny + 1
            ok &= np.array_equal(loc[jl - 1 + Hy], G[jg - 1 + mc.Hy])
        ok &= np.array_equal(G[mc.Hy], G[mc.Hy + mc.Ny]) and bool((loc > 0).all())      # face Ny + 1 is face 1
    # a north fold on y-slabs: only the last slab carries copy lists; in (i, j) relative to the slab's north edge they are the
    # global case's lists (rows within the global halo), plus the deeper rows of the slab's own halo; sources stay inside the slab
    fc = folded_case(24, 16 * world, H=7, substeps=K)
    fs = slab_of(fc, rank, world, Hy)
    ok &= (fs.fold is not None) == (rank == world - 1)
    if fs.fold is not None:
        for loc, (tg, sr) in fs.fold["maps"].items():
            sxp = fs.Nx + 2 * fs.Hx

            def rel(idx, H, Ny):   # (column, row above the north edge) of a linear parent index
                return idx % sxp, idx // sxp - (Ny - 1 + H)

            mine = {rel(t, Hy, fs.Ny): rel(q, Hy, fs.Ny) for t, q in zip(tg, sr)}
            gt, gs = fc.fold["maps"][loc]
            for t, q in zip(gt, gs):
                ok &= mine.get(rel(t, fc.Hy, fc.Ny)) == rel(q, fc.Hy, fc.Ny)
            ok &= all(-fs.Ny < r <= 0 for _, r in mine.values())          # every source is an interior row of this slab
            ok &= np.unique(tg).size == tg.size and not np.intersect1d(tg, sr).size
        m = fs.mask.reshape(-1)
        tg, sr = fs.fold["maps"][(0, 0)]
        ok &= np.array_equal(m[tg], m[sr])                                  # the slab's mask obeys the fold in its deeper halo too
    flag = torch.tensor([0 if ok else 1])
    dist.all_reduce(flag)
    if rank == 0:
        print("GLOO_SLABS_OK" if flag.item() == 0 else "GLOO_SLABS_BAD", flush=True)
    dist.destroy_process_group()
    sys.exit(int(flag.item() != 0))


if __name__ == "__main__":
    main()
