#!/bin/bash
# Multi-GPU check on one box: the N-GPU == 1-GPU tests and bench.py at N ranks.  usage: tools/gpu_multi.sh TAG N
cd "$(dirname "$0")/.."
TAG=${1:-x}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 2400 python -m pytest tests/test_multigpu.py -m gpu -q -rs > gpurun_out/multigpu_tests_${TAG}_n$N.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/multigpu_tests_${TAG}_n$N.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_${TAG}_n$N.json; tail -3 gpurun_out/bench_${TAG}_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_ref_n$N.json 2>/dev/null; tail -c 600 gpurun_out/bench_${TAG}_ref_n$N.json
