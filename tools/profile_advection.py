"""The advection kernels (a16 compute_tracer_tendencies!, a17 dynamic_time_step!) on one GPU, for ncu / quick timing:
    python tools/profile_advection.py [N] [reps]
Prints the average launch time of each and its achieved algorithmic bandwidth (48 B per cell each: DESIGN.md section 5)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as e; e.load_package()
import torch
from climaseaice_b200.driver import model_from_case
from climaseaice_b200.synthetic import anticyclone_case

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
case = anticyclone_case(N, substeps=2)
m = model_from_case(case)
m.update_state()
m.cache_current_fields()
for name, call in (("k_tracer_tendencies", m.compute_tracer_tendencies), ("k_dynamic_step", lambda: m.dynamic_time_step(1.0))):
    call()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(reps):
        call()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    print(f"{name}: {ms:.4f} ms per launch at {N}x{N}, {48 * N * N / ms / 1e6:.0f} GB/s of algorithmic traffic (48 B per cell)")
