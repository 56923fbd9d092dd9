#!/bin/bash
# A subset of the GPU tests by keyword.  usage: tools/run_subset.sh "KEYWORDS" [TAG]
cd "$(dirname "$0")/.."
K=${1:-fold}
TAG=${2:-subset}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -k "$K" > gpurun_out/subset_$TAG.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/subset_$TAG.log
