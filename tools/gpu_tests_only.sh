#!/bin/bash
# pytest -m gpu on one box, nothing else.  usage: tools/gpu_tests_only.sh TAG
cd "$(dirname "$0")/.."
TAG=${1:-x}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/gputests_$TAG.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
