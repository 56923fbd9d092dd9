"""BASELINE config 5 (lat-lon Arctic cap 4320 x 336, coupled) on the N GPUs of a torchrun launch, for several exchange intervals K
(halo 2K + 3 rows): python -m torch.distributed.run --nproc-per-node N tools/config5_k_sweep.py 4 8 12"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as e; e.load_package()
import torch
import torch.distributed as dist
from climaseaice_b200 import nccl_unique_id
from climaseaice_b200.driver import model_from_case
from climaseaice_b200.synthetic import arctic_cap_case, slab_of

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
dev = f"cuda:{local}"
c5 = arctic_cap_case(4320, 336, H=7, substeps=150, dt=600.0)
for K in [int(a) for a in sys.argv[1:]] or [4]:
    if 336 // world < 2 * K + 3:
        continue
    s5 = slab_of(c5, rank, world, 2 * K + 3)
    m5 = model_from_case(s5, solver_impl="auto", partition=(rank, world, K), device=dev)
    ids = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    m5.comm_init(ids[0])
    m5.time_step(c5.dt)
    torch.cuda.synchronize(); dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(3):
        m5.time_step(c5.dt)
    ev1.record()
    torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([ev0.elapsed_time(ev1) / 3], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"config 5 on {world} GPUs, K = {K:2d} (halo {2 * K + 3} rows, {s5.Ny} + {2 * (2 * K + 3)} rows per rank): {t.item():.2f} ms per time_step!, tiles {m5.fused_stats()[2]}", flush=True)
    m5.close()
    del m5
    torch.cuda.empty_cache()
dist.destroy_process_group()
