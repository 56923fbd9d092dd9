#!/bin/bash
# One GPU-box visit: parity tests, A/B bench, one ncu capture of the fused kernel.  usage: tools/gpu_round.sh TAG [notest]
cd "$(dirname "$0")/.."
TAG=${1:-x}
mkdir -p gpurun_out
if [ "$2" != "notest" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gputests_$TAG.log
fi
AB_STEPS=5 tools/ab_bench.sh base 2>&1 | tee gpurun_out/ab_$TAG.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_evp_substep_fused -s 20 -c 1 -f -o gpurun_out/fused_$TAG python tools/profile_case.py 4096 30 fused bounded > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
