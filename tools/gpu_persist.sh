#!/bin/bash
# The persistent small-grid kernel: parity tests (bounded by a timeout: a wrong grid barrier would spin), then config 1 with and without it
cd "$(dirname "$0")/.."
TAG=${1:-pers}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_c_abi.py -m gpu -x -q > gpurun_out/persist_tests_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/persist_tests_$TAG.log
CSI_PERSISTENT=0 timeout 300 tools/small_grid_ab.sh base 2>&1 | sed 's/^base/launch per substep/' | tee gpurun_out/persist_ab_$TAG.txt
timeout 300 tools/small_grid_ab.sh base 2>&1 | sed 's/^base/one cooperative launch/' | tee -a gpurun_out/persist_ab_$TAG.txt
