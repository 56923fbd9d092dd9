#!/bin/bash
# config 1 as shipped (128 x 128) under library variants: tools/small_grid_ab.sh variant.so ...
cd "$(dirname "$0")/.."
for v in "$@"; do
    if [ "$v" = "base" ]; then unset CSI_B200_LIB; else export CSI_B200_LIB="$PWD/$v"; fi
    python - "$v" <<'PY'
import sys
sys.path.insert(0, ".")
import __graft_entry__ as e; e.load_package()
import torch
from climaseaice_b200.driver import model_from_case
from climaseaice_b200.synthetic import anticyclone_case
for N in (128, 256, 512):
    case = anticyclone_case(N, noise=0.0)
    m = model_from_case(case)
    m.time_step(case.dt); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(5): m.time_step(case.dt)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 5
    print(f"{sys.argv[1]:24s} N={N:4d}: {ms:.3f} ms per time_step!, {ms*1e3/450:.2f} us per substep, tiles {m.fused_stats()[2]}")
    m.close()
PY
done
