#!/bin/bash
# A/B on one GPU box with a quick parity subset for every variant: tools/ab_round.sh TAG variant.so ...
cd "$(dirname "$0")/.."
TAG=$1; shift
for spec in "$@"; do
    v="${spec%%@*}"
    if [ "$v" = "base" ]; then unset CSI_B200_LIB; else export CSI_B200_LIB="$PWD/$v"; fi
    timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/par_${TAG}_$(basename $v .so).log 2>&1; echo "parity $v rc=$? $(tail -1 gpurun_out/par_${TAG}_$(basename $v .so).log)"
done
unset CSI_B200_LIB
AB_STEPS=5 tools/ab_bench.sh "$@" 2>&1 | tee gpurun_out/ab_$TAG.txt
