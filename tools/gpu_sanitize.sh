#!/bin/bash
# compute-sanitizer memcheck and racecheck over tools/sanitize_case.py, then the new mid-size test
cd "$(dirname "$0")/.."
TAG=${1:-san}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "many_tiles" 2>&1 | tail -3
{ echo "--- memcheck ---"; timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_case.py 2>&1 | tail -18
  echo "--- racecheck ---"; timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_case.py 2>&1 | tail -18; } > gpurun_out/sanitizer_$TAG.txt
grep "SUMMARY\|ok" gpurun_out/sanitizer_$TAG.txt | tail -30
