#!/bin/bash
# Quick timings of the fused kernel's three metric modes (regular, per-row, two-dimensional) and of the general kernels on the latter
cd "$(dirname "$0")/.."
TAG=${1:-modes}
mkdir -p gpurun_out
{ for k in bounded latlon curvilinear; do python tools/profile_case.py 2048 150 fused $k; done; python tools/profile_case.py 2048 150 unfused curvilinear; } 2>&1 | grep "cell-updates" | tee gpurun_out/${TAG}_timing.txt
