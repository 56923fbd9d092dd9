#!/bin/bash
# the default bench.py run on one box, nothing else.  usage: tools/gpu_bench_only.sh TAG
cd "$(dirname "$0")/.."
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err; echo "bench rc=$?"; tail -c 1200 gpurun_out/bench_${TAG}_n1.json; tail -3 gpurun_out/bench_${TAG}_n1.err
