#!/bin/bash
# Round profiles on one box: ncu --set full of the fused kernel and of the two advection kernels, and the launch list of bench.py
cd "$(dirname "$0")/.."
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_evp_substep_fused -s 20 -c 1 -f -o gpurun_out/${TAG}_fused python tools/profile_case.py 4096 30 fused bounded > gpurun_out/${TAG}_ncu_fused.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_fused.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tracer_tendencies -s 2 -c 1 -f -o gpurun_out/${TAG}_tendencies python tools/profile_advection.py 4096 4 > gpurun_out/${TAG}_ncu_adv1.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_adv1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dynamic_step -s 2 -c 1 -f -o gpurun_out/${TAG}_dynstep python tools/profile_advection.py 4096 4 > gpurun_out/${TAG}_ncu_adv2.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_adv2.log
python tools/profile_advection.py 4096 20 | tee gpurun_out/${TAG}_advection_timing.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-configs > gpurun_out/${TAG}_launches_bench.json 2>/dev/null; wc -l gpurun_out/${TAG}_launches_bench.csv
