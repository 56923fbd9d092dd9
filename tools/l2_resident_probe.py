"""How much faster is the fused substep kernel when its working set stays in L2?  Doubly periodic cases whose tile counts are whole
waves of 444 CTAs: 1110 x 336 (2 waves, 72 MB of planes: L2-resident), 1110 x 672 (4 waves), 4440 x 1344 (32 waves, HBM regime).
    python tools/l2_resident_probe.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as e; e.load_package()
import torch
from climaseaice_b200.driver import model_from_case
from climaseaice_b200.synthetic import periodic_case
for nx, ny in ((1110, 336), (1110, 672), (2220, 672), (4440, 1344)):
    case = periodic_case(nx, Ny=ny, substeps=150, aice="ones")
    m = model_from_case(case, solver_impl="fused")
    m.update_state()
    m.time_step_momentum(case.dt, 150)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(4):
        m.time_step_momentum(case.dt, 150)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 4
    print(f"{nx}x{ny}: {ms / 150 * 1e3:.1f} us per substep, {nx * ny * 150 / ms / 1e6:.3f} G cell-updates/s, tiles {m.fused_stats()[2]}, planes {24 * nx * ny * 8 / 1e6:.0f} MB")
    m.close()
