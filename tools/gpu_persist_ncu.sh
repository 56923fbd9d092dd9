#!/bin/bash
# the new test, then ncu --set full of the persistent kernel on BASELINE config 1 as shipped (128 x 128)
cd "$(dirname "$0")/.."
TAG=${1:-pers}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cooperative" 2>&1 | tail -5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_evp_substeps_persistent -s 3 -c 1 -f -o gpurun_out/${TAG}_persistent python tools/profile_case.py 128 150 fused bounded > gpurun_out/${TAG}_ncu.log 2>&1; tail -2 gpurun_out/${TAG}_ncu.log
