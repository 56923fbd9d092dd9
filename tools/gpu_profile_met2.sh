#!/bin/bash
# ncu --set full of the fused kernel's two-dimensional-metric instantiation, and quick timings of the three metric modes
cd "$(dirname "$0")/.."
TAG=${1:-met2}
mkdir -p gpurun_out
for k in bounded latlon curvilinear; do python tools/profile_case.py 2048 30 fused $k; done 2>&1 | grep "cell-updates" | tee gpurun_out/${TAG}_timing.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_evp_substep_fused -s 20 -c 1 -f -o gpurun_out/${TAG}_fused python tools/profile_case.py 2048 30 fused curvilinear > gpurun_out/${TAG}_ncu.log 2>&1; tail -1 gpurun_out/${TAG}_ncu.log
