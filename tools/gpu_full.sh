#!/bin/bash
# Full check on one box: GPU tests, default bench (N = 1), reference arm.  usage: tools/gpu_full.sh TAG
cd "$(dirname "$0")/.."
TAG=${1:-x}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gputests_$TAG.log
timeout 1200 python bench.py > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_${TAG}_n1.json; tail -3 gpurun_out/bench_${TAG}_n1.err
