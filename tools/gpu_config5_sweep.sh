#!/bin/bash
# usage: tools/gpu_config5_sweep.sh N K1 K2 ...
cd "$(dirname "$0")/.."
N=$1; shift
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/config5_k_sweep.py "$@" 2>&1 | grep "config 5" | tee gpurun_out/config5_sweep_n$N.txt
