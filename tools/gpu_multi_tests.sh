#!/bin/bash
# Multi-GPU tests only (N-GPU == 1-GPU).  usage: tools/gpu_multi_tests.sh TAG [pytest -k expression]
cd "$(dirname "$0")/.."
TAG=${1:-x}
KEXPR=${2:-gpus}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 2400 python -m pytest tests/test_multigpu.py -m gpu -q -rs -x -k "$KEXPR" > gpurun_out/multigpu_tests_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/multigpu_tests_${TAG}.log
