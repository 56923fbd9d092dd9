#!/bin/bash
# bench.py at N ranks on one box: tools/gpu_bench_n.sh TAG N
cd "$(dirname "$0")/.."
TAG=${1:-x}; N=${2:-2}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; echo "bench rc=$?"; tail -c 1800 gpurun_out/bench_${TAG}_n$N.json; tail -3 gpurun_out/bench_${TAG}_n$N.err
