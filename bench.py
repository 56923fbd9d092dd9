#!/usr/bin/env python
"""bench.py -- EVP substep cell-updates/s (Float64) of the B200 hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A "step" is one `time_step_momentum!` (reset + initialize_rheology + `substeps` EVP substeps +
finalize) over one synthetic batch.  N = 1: BASELINE config 2, the anticyclone case scaled to
4096 x 4096 (Bounded x Bounded).  N > 1: BASELINE config 3, doubly periodic, y-slabs of
16384 x 2048 per GPU (16384^2 at N = 8), NCCL halo exchange every K substeps; weak scaling.
`value`  : cell-updates/s with inputs resident in HBM, CUDA events, max over ranks.
`e2e`    : the same metric through the host-buffer C-ABI call (pinned host arrays, H2D + D2H inside).
`roofline`: dominant kernel, algorithmic bytes (144 B per cell-update fused / see DESIGN.md) over its average launch
            duration in the timed region (CUDA events), against MEASURED_PEAKS.json hbm_gbs; `*_alone`: the kernel in a burst.
`cpu_baseline`: the CPU oracle (C restatement of the reference's KernelAbstractions-CPU path; the
            Julia original cannot run in this image) on all host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "EVP substep cell-updates/s (Float64)"
UNIT = "cell-updates/s"
SUBSTEPS = 150
DT_STAGE = 120.0
BYTES_PER_CELL_FUSED = 144      # r/w u,v,s11,s22,s12 + read h,aice,un,vn,tau_x,tau_y,ue,ve
BYTES_PER_CELL_STRESS = 120     # unfused stress kernel: read u,v,P,h,aice,s11,s22,s12; write s11,s22,s12,zeta_c,zeta_f,Delta,alpha


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text()).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(n, steps, warmup, threads=None, periodic=False):
    """Times the oracle's time_step_momentum! (150 substeps) on an n x n sample of the workload (anticyclone / periodic)."""
    import __graft_entry__ as entry
    entry.load_package()
    from climaseaice_b200.synthetic import anticyclone_case, periodic_case
    from oracle import oracle as O
    from tests.helpers import oracle_from_case
    cores = O.set_threads(threads or (os.cpu_count() or 1))
    case = periodic_case(n, substeps=SUBSTEPS, aice="mixed") if periodic else anticyclone_case(n, substeps=SUBSTEPS)
    o = oracle_from_case(case)
    o.update_state()
    for _ in range(warmup):
        o.time_step_momentum(DT_STAGE, SUBSTEPS)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.time_step_momentum(DT_STAGE, SUBSTEPS)
    dt = time.perf_counter() - t0
    return n * n * SUBSTEPS * steps / dt, cores, dt / steps


def reference_sample_size(steps, warmup):
    budget_cells = 1.0e7 * 150.0 / max(1, steps + warmup) / SUBSTEPS  # ~150 s of CPU work at ~1e7 cell-updates/s
    n = 128
    while (2 * n) ** 2 <= budget_cells and 2 * n <= 2048:
        n *= 2
    return n


def run_reference(args):
    """The reference arm: the reference's CPU implementation of the path (its C restatement; Julia is not installable in
    this image) on all host cores, on a bounded sample of this arm's workload.  `config.sample_grid` is the grid that ran."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = reference_sample_size(args.steps, args.warmup)
    periodic = args.gpus > 1        # N > 1 times the doubly periodic case (config 3), N = 1 the Bounded anticyclone (config 2)
    rate, cores, per_step = cpu_reference_rate(n, args.steps, args.warmup, periodic=periodic)
    kind = "doubly periodic" if periodic else "anticyclone (Bounded x Bounded)"
    sample = f"{kind} {n}x{n}, {SUBSTEPS} substeps per step (bounded sample of the workload named in config.workload)"
    cfg = workload_config(args.gpus)
    cfg["sample_grid"] = [n, n]
    cfg["note"] = "config.grid names the GPU arm's workload; the CPU arm ran sample_grid (same physics, same substep count)"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "C restatement of the reference's KernelAbstractions-CPU path (OpenMP); the Julia original is not runnable in this image"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(ngpus, Rx=1):
    if ngpus > 1 and Rx > 1:
        Ry = ngpus // Rx
        return {"workload": f"doubly periodic EVP on 16384x{2048 * ngpus} (BASELINE config 3 at 8 GPUs = 16384^2), {Rx}x{Ry} blocks of "
                            f"{16384 // Rx}x{2048 * ngpus // Ry} per GPU, NCCL halo exchange (packed west/east strips, zero-copy rows) every 4 substeps, "
                            "150 substeps per step",
                "grid": [16384, 2048 * ngpus], "per_gpu": [16384 // Rx, 2048 * ngpus // Ry], "partition": [Rx, Ry], "substeps": SUBSTEPS,
                "exchange_every": 4, "l2_policy": "inputs larger than L2 (4.9 GB working set per GPU vs 126 MB)"}
    if ngpus == 1:
        return {"workload": "anticyclone EVP benchmark scaled to 4096x4096 (BASELINE config 2): Bounded x Bounded, H=7, dx=4 km, "
                            "FPlane f=1e-4, wind-stress arrays + SemiImplicitStress ocean drag, 150 substeps per step",
                "grid": [4096, 4096], "substeps": SUBSTEPS, "l2_policy": "inputs larger than L2 (2.4 GB working set vs 126 MB)"}
    return {"workload": f"doubly periodic EVP on 16384x{2048 * ngpus} (BASELINE config 3 at 8 GPUs = 16384^2), y-slabs of 16384x2048 per GPU, "
                        "NCCL halo exchange every 4 substeps, 150 substeps per step",
            "grid": [16384, 2048 * ngpus], "per_gpu": [16384, 2048], "substeps": SUBSTEPS, "exchange_every": 4,
            "l2_policy": "inputs larger than L2 (4.9 GB working set per GPU vs 126 MB)"}


# ------------------------------------------------------------------------------------------------
def multi_gpu_parity(rank, world, local, Rx, dist):
    """N GPUs == 1 GPU, bit for bit, before anything is timed (the analogue of test/distributed_tests_utils.jl:40-88): a small
    global doubly periodic case is stepped twice (RK3, fused solver) as world blocks with NCCL halo exchange every K substeps
    and, by every rank on its own GPU, as one domain; each rank compares its block of u, v, h, aice, sigma."""
    import torch
    from climaseaice_b200 import nccl_unique_id
    from climaseaice_b200.driver import model_from_case
    from climaseaice_b200.synthetic import block_of, periodic_case, slab_of
    K, Ry = 3, world // Rx
    Hy = 2 * K + 3
    case = periodic_case(96 * Rx, Ny=32 * Ry, substeps=10, aice="mixed")
    dev = f"cuda:{local}"
    if Rx > 1:
        sl = block_of(case, rank, Rx, Ry, Hy, Hy)
        m = model_from_case(sl, solver_impl="fused", partition=(rank, world, K, Rx), device=dev)
    else:
        sl = slab_of(case, rank, world, Hy)
        m = model_from_case(sl, solver_impl="fused", partition=(rank, world, K), device=dev)
    ids = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    m.comm_init(ids[0])
    ref = model_from_case(case, solver_impl="fused", device=dev)
    for _ in range(2):
        m.time_step(case.dt)
        ref.time_step(case.dt)
    torch.cuda.synchronize()
    rx, ry = rank % Rx, rank // Rx
    bad = 0
    for n in ("u", "v", "h", "a", "s11", "s22", "s12"):
        mine = m.all_fields()[n].parent[Hy:Hy + sl.Ny, sl.Hx:sl.Hx + sl.Nx]
        glob = ref.all_fields()[n].parent[case.Hy + ry * sl.Ny:case.Hy + (ry + 1) * sl.Ny, case.Hx + rx * sl.Nx:case.Hx + (rx + 1) * sl.Nx]
        if not torch.equal(mine, glob):
            bad += 1
            print(f"bench.py: rank {rank}: multi-GPU parity check: field {n} differs from the 1-GPU run", file=sys.stderr, flush=True)
    used_fused = m.fused_stats()[2] > 0
    flag = torch.tensor([bad + (0 if used_fused else 1)], device=dev)
    dist.all_reduce(flag)
    m.close(); ref.close()
    return int(flag.item()) == 0, f"{case.Nx}x{case.Ny} doubly periodic, 2 time steps (RK3, 10 substeps, exchange every {K}), fused solver, {Rx}x{Ry} blocks; u v h aice s11 s22 s12"


def timed_momentum_steps(model, steps, warmup, barrier, local):
    import torch
    for _ in range(warmup):
        model.time_step_momentum(DT_STAGE, SUBSTEPS)
    barrier()
    launches0 = model.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        ev0.record()
        for _ in range(steps):
            model.time_step_momentum(DT_STAGE, SUBSTEPS)
        ev1.record()
        barrier()
    return ev0.elapsed_time(ev1), model.launch_count - launches0, clk


def extra_configuration_legs(dev, local, steps=2):
    """The other named BASELINE configurations, on one GPU, each a bounded run (beside the headline, never instead of it)."""
    import torch
    from climaseaice_b200.driver import model_from_case
    from climaseaice_b200.synthetic import anticyclone_case, arctic_cap_case, coastline_case
    out = {}

    def time_full_steps(case, n, solver="auto"):
        m = model_from_case(case, solver_impl=solver, device=dev)
        m.time_step(case.dt)                     # warm-up; the first step also runs update_state!
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = m.launch_count
        ev0.record()
        for _ in range(n):
            m.time_step(case.dt)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / n
        st = m.fused_stats()
        launches = (m.launch_count - l0) // n
        m.close()
        return ms, st, launches

    # config 1: examples/ice_advected_by_anticyclone.jl as shipped (128^2, RK3, WENO7, 150 substeps): the launch-latency regime
    c1 = anticyclone_case(128, noise=0.0)
    ms, st, nl = time_full_steps(c1, 5)
    out["config1_anticyclone_128_as_shipped"] = {
        "ms_per_time_step": ms, "us_per_substep": ms * 1e3 / (3 * SUBSTEPS), "launches_per_time_step": nl,
        "cell_updates_per_s": 128 * 128 * 3 * SUBSTEPS / (ms * 1e-3), "fused_stats": list(st),
        "note": "one time_step! = 3 RK stages x (WENO7 tendencies + 150 substeps + h/aice update); 16 k cells do not fill 148 SMs; "
                "the substeps of a stage but the last run as one cooperative launch with a grid-wide barrier (CSI_PERSISTENT=0: one launch per substep)"}
    # config 4: ice_advected_on_coastline with the immersed land mask, 8192 x 4096 (2 Ny x Ny as the example), momentum only
    c4 = coastline_case(Ny=4096, substeps=SUBSTEPS)
    m = model_from_case(c4, solver_impl="auto", device=dev)
    m.update_state()
    m.time_step_momentum(DT_STAGE, SUBSTEPS)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        m.time_step_momentum(DT_STAGE, SUBSTEPS)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    out["config4_coastline_8192x4096_masked"] = {
        "ms_per_step": ms, "cell_updates_per_s": c4.Nx * c4.Ny * SUBSTEPS / (ms * 1e-3), "fused_stats": list(m.fused_stats()),
        "land_fraction": float(c4.mask[c4.Hy:-c4.Hy, c4.Hx:-c4.Hx].mean()),
        "note": "Periodic x Bounded, immersed triangular coast, linear immersed drag, uniform wind, ocean at rest; one time_step_momentum! of 150 substeps"}
    m.close()
    del m
    torch.cuda.empty_cache()
    # config 5 on one GPU: coupled thermodynamics + dynamics + advection on the 1/12-degree lat-lon cap (4320 x 336)
    c5 = arctic_cap_case(4320, 336, H=7, substeps=SUBSTEPS, dt=600.0)
    ms, st, nl = time_full_steps(c5, steps)
    out["config5_arctic_cap_4320x336_one_gpu"] = {
        "ms_per_time_step": ms, "cell_updates_per_s": c5.Nx * c5.Ny * 3 * SUBSTEPS / (ms * 1e-3), "launches_per_time_step": nl,
        "fused_stats": list(st),
        "note": "lat-lon metrics, HydrostaticSphericalCoriolis, slab thermodynamics, RK3; the 8-GPU slabs of this case are 4320 x 42 per rank"}
    torch.cuda.empty_cache()
    # beyond BASELINE's list: the grids of the "next" row (f2) -- two-dimensional metrics (fused tile kernel reading per-node metric
    # planes, beside the general kernels) and a tripolar-like mesh with a north fold (general kernels), momentum only
    from climaseaice_b200.synthetic import curvilinear_case, folded_case

    def time_momentum(case, solver, n):
        m = model_from_case(case, solver_impl=solver, device=dev)
        m.update_state()
        m.time_step_momentum(DT_STAGE, SUBSTEPS)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(n):
            m.time_step_momentum(DT_STAGE, SUBSTEPS)
        ev1.record()
        torch.cuda.synchronize()
        ms, st = ev0.elapsed_time(ev1) / n, m.fused_stats()
        m.close()
        del m
        torch.cuda.empty_cache()
        return ms, st

    cc = curvilinear_case(2048, 2048, H=7, substeps=SUBSTEPS, topology=("Periodic", "Bounded"))
    msf, stf = time_momentum(cc, "fused", steps)
    msu, _ = time_momentum(cc, "unfused", 1)
    out["extra_curvilinear_2048_two_dimensional_metrics"] = {
        "fused_ms_per_step": msf, "fused_cell_updates_per_s": cc.Nx * cc.Ny * SUBSTEPS / (msf * 1e-3), "fused_stats": list(stf),
        "general_kernels_ms_per_step": msu, "general_kernels_cell_updates_per_s": cc.Nx * cc.Ny * SUBSTEPS / (msu * 1e-3),
        "note": "orthogonal curvilinear mesh (spacings vary +-25 % along both axes), Periodic x Bounded; one time_step_momentum! of 150 substeps"}
    del cc
    cf = folded_case(2048, 1024, H=7, substeps=SUBSTEPS)
    msf, stf = time_momentum(cf, "fused", steps)
    msu, _ = time_momentum(cf, "unfused", 1)
    out["extra_tripolar_like_2048x1024_north_fold"] = {
        "fused_ms_per_step": msf, "fused_cell_updates_per_s": cf.Nx * cf.Ny * SUBSTEPS / (msf * 1e-3), "fused_stats": list(stf),
        "general_kernels_ms_per_step": msu, "general_kernels_cell_updates_per_s": cf.Nx * cf.Ny * SUBSTEPS / (msu * 1e-3),
        "note": "two-dimensional metrics, zonally periodic, north fold (copy lists), an island at the fold; the tile rows next to the fold run "
                "the substep as two launches with the fold fill of the first velocity between them (DESIGN.md section 7)"}
    return out


def fp64_roofline(local, cell_updates_per_s):
    """The bound that actually binds: FP64 instruction issue.  achieved = FP64 thread-instructions per cell-update of this
    build (ncu, profiles/dominant_kernel_traffic.json) x cell-updates/s; peak = FP64 FMA instructions/s measured on this
    device just now (csi_measure_fp64_rate), sustained under the power cap and as a burst."""
    import ctypes as C
    from climaseaice_b200 import _lib as L
    p = ROOT / "profiles" / "dominant_kernel_traffic.json"
    per_cell = None
    if p.exists():
        per_cell = json.loads(p.read_text()).get("k_evp_substep_fused", {}).get("fp64_thread_instr_per_cell_update")
    burst, sust = C.c_double(), C.c_double()
    rc = L.lib().csi_measure_fp64_rate(local, 2.0, C.byref(burst), C.byref(sust))
    if rc != 0 or not per_cell:
        return None
    achieved = per_cell * cell_updates_per_s
    return {"bound": "fp64", "unit": "FP64 thread-instructions/s", "achieved": achieved, "peak": sust.value, "frac": achieved / sust.value,
            "peak_burst": burst.value, "frac_of_burst": achieved / burst.value, "fp64_instr_per_cell_update": per_cell,
            "peak_source": "csi_measure_fp64_rate on this device in this run: 8 independent DFMA chains per thread, 2 s back to back, mean of the second half (power-capped clock)"}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    entry.load_package()
    from climaseaice_b200 import _lib as L, nccl_unique_id
    from climaseaice_b200.driver import HostStepper, model_from_case
    from climaseaice_b200.synthetic import anticyclone_case, periodic_case, slab_of

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    ngpus = world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    nx, ny = (args.nx or 4096, args.ny or 4096) if ngpus == 1 else (args.nx or 16384, args.ny or 2048)
    K = 4
    Rx = max(1, args.partition_x)
    parity = None
    if ngpus > 1:
        if ngpus % Rx:
            raise SystemExit("bench.py: --partition-x must divide the number of GPUs")
        ok, what = multi_gpu_parity(rank, world, local, Rx, dist)
        parity = {"result": "bitwise" if ok else "MISMATCH", "case": what}
        if not ok:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "error": "multi-GPU run differs from the 1-GPU run", "multi_gpu_parity": parity}), flush=True)
            dist.destroy_process_group()
            return 1
    if ngpus == 1:
        case = anticyclone_case(nx) if not args.periodic else periodic_case(nx, Ny=ny)
        model = model_from_case(case, solver_impl=args.solver, device=dev)
        part = None
    else:
        # each rank builds only its own slab of the global periodic case (same seed => consistent fields)
        Hy = 2 * K + 3
        Ry = ngpus // Rx
        case = periodic_slab_case(nx // Rx, ny * ngpus // Ry, rank // Rx, Ry, Hy, rx=rank % Rx, Rx=Rx)
        part = (rank, ngpus, K, Rx) if Rx > 1 else (rank, ngpus, K)
        model = model_from_case(case, solver_impl=args.solver, partition=part, device=dev)
        ids = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        model.comm_init(ids[0])
    cells = case.Nx * case.Ny
    model.update_state()
    torch.cuda.synchronize()

    ms, launches, clk = timed_momentum_steps(model, args.steps, args.warmup, barrier, local)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = cells * ngpus * SUBSTEPS / (ms_per_step * 1e-3)
    stats = model.fused_stats()

    # dominant kernel: its average launch duration over the timed region (CUDA events on its stream).  The fused solver
    # launches it once per substep, so that is the region's time over the launches -- an upper bound, the region also holds
    # the stage's pack / prep / unpack / init kernels (about 1 %) and, for N > 1, the halo exchange.  The same kernel launched
    # alone in a short burst (higher clock, no power cap yet) is reported beside it.
    kern_alone_ms, kern_name, kern_bytes = time_dominant_kernel(model, cells)
    if kern_name == "k_evp_substep_fused":
        kern_ms, kern_src = ms_per_step / SUBSTEPS, "timed region / launches (includes the stage's pack, prep, unpack and init kernels, about 1 %)"
    else:
        kern_ms, kern_src = kern_alone_ms, "kernel launched alone between CUDA events (several kernels per substep)"
    peak, peak_src = peaks()
    achieved = kern_bytes / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": kern_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src, "traffic": traffic_from_profile(kern_name, cells), "launch_ms": kern_ms, "launch_ms_source": kern_src,
                "algorithmic_bytes_per_launch": kern_bytes,
                "launch_ms_alone": kern_alone_ms, "achieved_alone": kern_bytes / (kern_alone_ms * 1e-3) / 1e9 if kern_alone_ms > 0 else None,
                "note": "the kernel is bound by FP64 instruction issue under the board's power cap, not by HBM: see roofline.fp64; temporal blocking "
                        "across substeps (north_star: optional) is not used -- it trades HBM bytes, which are not the bound, for redundant FP64 work, which is"}
    if rank == 0 and kern_name == "k_evp_substep_fused":
        roofline["fp64"] = fp64_roofline(local, cells * SUBSTEPS / (ms_per_step * 1e-3))

    # full model step (3 RK stages incl. advection), reported beside the headline
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    model.time_step(case.dt)
    ev1.record()
    barrier()
    full_ms = ev0.elapsed_time(ev1)
    model.close()
    del model
    torch.cuda.empty_cache()

    # end to end through the host-buffer entry point: every rank's pinned host block (halos included) goes up, the results
    # come down, inside the timed call; wall clock between barriers, max over ranks
    e2e = None
    if not args.no_e2e:
        uid = None
        if world > 1:
            ids = [nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            uid = ids[0]
        hs = HostStepper(case, solver_impl=args.solver, device_index=local, partition=part, unique_id=uid)
        hs.evp_substeps(DT_STAGE, SUBSTEPS)  # warm-up (allocates the device mirrors)
        n_e2e = max(1, min(args.steps, 3))
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            hs.evp_substeps(DT_STAGE, SUBSTEPS)
        barrier()
        wall = (time.perf_counter() - t0) / n_e2e
        if world > 1:
            t = torch.tensor([wall], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wall = float(t.item())
        h2d, d2h = hs.last_transfer_bytes()
        e2e = {"value": cells * ngpus * SUBSTEPS / wall, "unit": UNIT, "h2d_bytes_per_step": h2d * ngpus, "d2h_bytes_per_step": d2h * ngpus,
               "ms_per_step": wall * 1e3, "call": "csi_evp_substeps_host (pinned host arrays; per-rank blocks with halos when N > 1)"}
        hs.model.close()
        del hs
        torch.cuda.empty_cache()

    # BASELINE config 5 on the N GPUs of this run: coupled thermodynamics + dynamics + advection on the 1/12-degree lat-lon cap,
    # y-slabs of 4320 x (336 / N) with NCCL halo exchange (beside the headline, never instead of it)
    config5 = None
    if world > 1 and not args.no_configs and 336 % world == 0:
        from climaseaice_b200.synthetic import arctic_cap_case
        c5 = arctic_cap_case(4320, 336, H=7, substeps=SUBSTEPS, dt=600.0)
        # thin slabs: fewer, deeper exchanges (measured on 4 GPUs, 84 rows per rank: 25.96 ms at K = 4, 23.85 at K = 8; the 42-row slabs of
        # 8 GPUs were measured at K = 4 only and keep it)
        K5 = 8 if 336 // world == 84 else K
        s5 = slab_of(c5, rank, world, 2 * K5 + 3)
        m5 = model_from_case(s5, solver_impl=args.solver, partition=(rank, world, K5), device=dev)
        ids = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        m5.comm_init(ids[0])
        m5.time_step(c5.dt)   # warm-up; the first step also runs update_state!
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(2):
            m5.time_step(c5.dt)
        ev1.record()
        barrier()
        t5 = torch.tensor([ev0.elapsed_time(ev1) / 2], device=dev, dtype=torch.float64)
        dist.all_reduce(t5, op=dist.ReduceOp.MAX)
        st5 = m5.fused_stats()
        config5 = {"ms_per_time_step": float(t5.item()), "cell_updates_per_s": c5.Nx * c5.Ny * 3 * SUBSTEPS / (float(t5.item()) * 1e-3),
                   "per_gpu": [s5.Nx, s5.Ny], "exchange_every": K5, "fused_stats": list(st5),
                   "note": "one time_step! = 3 RK stages x (WENO7 tendencies + 150 substeps + h/aice update + slab thermodynamics); lat-lon metrics, "
                           "HydrostaticSphericalCoriolis; slabs this thin are bound by the latency of a tile pass and the exchange, not by throughput"}
        m5.close()
        del m5
        torch.cuda.empty_cache()

    # weak-scaling baseline: the same per-GPU block (16384 x 2048, doubly periodic) on ONE GPU of this box in this run
    weak1 = None
    if world > 1 and rank == 0 and not args.no_weak_baseline:
        c1 = periodic_slab_case(nx // Rx, ny * ngpus // (ngpus // Rx), 0, 1, 7)
        m1 = model_from_case(c1, solver_impl=args.solver, device=dev)
        m1.update_state()
        ms1, _, _ = timed_momentum_steps(m1, min(args.steps, 5), 3, torch.cuda.synchronize, local)
        ms1 /= min(args.steps, 5)
        weak1 = {"value": c1.Nx * c1.Ny * SUBSTEPS / (ms1 * 1e-3), "unit": UNIT, "ms_per_step": ms1, "grid": [c1.Nx, c1.Ny],
                 "note": "the per-GPU block of this run as one doubly periodic domain on one GPU (rank 0, others idle): like-for-like weak-scaling baseline"}
        m1.close()
    if world > 1:
        dist.barrier()

    configs = None
    if ngpus == 1 and rank == 0 and not args.no_configs and not (args.nx or args.periodic):
        try:
            configs = extra_configuration_legs(dev, local)
        except Exception as exc:   # the legs stand beside the headline: a failure there is reported, it never costs the line
            configs = {"error": f"{type(exc).__name__}: {exc}"}

    if rank == 0:
        cpu = None
        if ngpus == 1 and not args.no_cpu:
            n = args.cpu_n
            rate, cores, per_step = cpu_reference_rate(n, 1, 1)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"anticyclone {n}x{n}, one time_step_momentum! of {SUBSTEPS} substeps after one warm-up step ({per_step:.1f} s)",
                   "note": "C restatement of the reference's KernelAbstractions-CPU path; Julia original not runnable here"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ngpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(ngpus, Rx) if not (args.nx or args.periodic) else {"workload": f"custom {case.name} {case.Nx}x{case.Ny} per GPU"},
            "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "solver": args.solver,
            "fused_stats": {"inputs_failed_validation": int(stats[0]), "tile_passes_redone_ieee": int(stats[1]), "tiles_per_substep": int(stats[2]),
                            "note": "last timed step, rank 0; a non-zero first entry or a large second one means the 3x slower IEEE pass ran"},
            "full_time_step": {"ms": full_ms, "cell_updates_per_s": cells * ngpus * 3 * SUBSTEPS / (full_ms * 1e-3),
                               "note": "one time_step! = 3 RK stages x (WENO7 tendencies + 150 substeps + h/aice update)"},
        }
        if parity:
            line["multi_gpu_parity"] = parity["result"]
            line["multi_gpu_parity_case"] = parity["case"]
        if weak1:
            line["weak_baseline_1gpu"] = weak1
        if configs:
            line["configs"] = configs
        if config5:
            line["configs"] = {f"config5_arctic_cap_4320x336_on_{ngpus}_gpus": config5}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def periodic_slab_case(nx, ny_local, rank, nranks, Hy, rx=0, Rx=1):
    """Rank-local slab (row `rank` of `nranks`; with Rx > 1 the block rx of that row, `nx` columns wide, halo Hy on both axes)
    of the global doubly periodic case, generated without materialising the global arrays."""
    import numpy as np
    from climaseaice_b200.synthetic import Case, LOC
    Ny = ny_local * nranks
    c = Case(f"periodic-slab{rank}.{rx}", nx, ny_local, Hy if Rx > 1 else 7, Hy, ("Periodic", "Periodic"), nx * 4000.0, ny_local * 4000.0)
    tp = 2 * np.pi
    Lx, Ly = nx * Rx * 4000.0, Ny * 4000.0
    rng = np.random.default_rng(20260417 + rank * Rx + rx)

    def nodes(loc):
        sy, sx = c.parent_shape(loc)
        i = np.arange(sx) - c.Hx + 1 + rx * nx
        j = np.arange(sy) - c.Hy + 1 + rank * ny_local
        x = ((i - 1) if loc[0] else (i - 0.5)) * 4000.0
        y = ((j - 1) if loc[1] else (j - 0.5)) * 4000.0
        return x[None, :], y[:, None]

    x, y = nodes(LOC["h"])
    h = 0.3 + 0.005 * (np.sin(3 * tp * x / Lx) + np.sin(2 * tp * y / Ly)) + 0 * x
    h = h + 1e-3 * rng.uniform(-1, 1, h.shape)
    a = 0.9 + 0.1 * rng.uniform(0, 1, h.shape)
    xu, yu = nodes(LOC["u"])
    xv, yv = nodes(LOC["v"])
    f = dict(h=h, a=a, u=0.05 * np.sin(tp * yu / Ly) * np.cos(tp * xu / Lx), v=-0.05 * np.sin(tp * xv / Lx) * np.cos(tp * yv / Ly),
             ue=0.01 * np.sin(tp * yu / Ly) + 0 * xu, ve=0.01 * np.sin(tp * xv / Lx) + 0 * yv,
             top_x=0.1 * np.sin(tp * yu / Ly) + 0 * xu, top_y=0.1 * np.cos(tp * xv / Lx) + 0 * yv)
    c.fields = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in f.items()}
    # the random parts of h, aice differ between ranks in the overlapping halos; the first exchange
    # (update_state!) makes them consistent before anything is timed
    return c


def time_dominant_kernel(model, cells, reps=20):
    """Average duration of the dominant kernel launched alone between two CUDA events."""
    import ctypes as C
    from climaseaice_b200 import _lib as L
    lib = L.lib()
    if not hasattr(lib, "csi_time_dominant_kernel"):
        return float("nan"), "n/a", 0
    f = model.csi_fields()
    ms = C.c_double()
    name = C.create_string_buffer(64)
    bpc = C.c_int32()
    lib.csi_time_dominant_kernel.argtypes = [C.c_void_p, C.POINTER(L.csi_fields), C.c_double, C.c_int32, C.POINTER(C.c_double), C.c_char_p,
                                             C.POINTER(C.c_int32), C.c_void_p]
    L.check(lib.csi_time_dominant_kernel(model._handle, C.byref(f), DT_STAGE, reps, C.byref(ms), name, C.byref(bpc), model._stream()), model._handle)
    return ms.value, name.value.decode(), bpc.value * cells


def traffic_from_profile(kernel_name, cells):
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/) of a launch over the same number
    of cells, or None (a capture of another grid size says nothing about this run)."""
    p = ROOT / "profiles" / "dominant_kernel_traffic.json"
    if p.exists():
        d = json.loads(p.read_text()).get(kernel_name, {})
        return d.get("dram_bytes_per_launch_by_cells", {}).get(str(int(cells)))
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--solver", default="auto", choices=["auto", "unfused", "fused"])
    ap.add_argument("--nx", type=int, default=0)
    ap.add_argument("--ny", type=int, default=0)
    ap.add_argument("--periodic", action="store_true")
    ap.add_argument("--partition-x", type=int, default=1, help="N > 1: Rx of an Rx x (N / Rx) block partition (default 1 = y-slabs)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="N = 1: skip the extra legs for BASELINE configs 1, 4, 5")
    ap.add_argument("--no-weak-baseline", action="store_true", help="N > 1: skip timing the per-GPU block on one GPU")
    ap.add_argument("--cpu-n", type=int, default=2048)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
