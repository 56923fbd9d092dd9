#!/bin/bash
# A/B of kernel variants on one GPU box: tools/ab_bench.sh [variant.so ...]   ("base" = the in-tree library)
# Prints value (cell-updates/s), the fused kernel's launch ms and the SM clock for each variant.
cd "$(dirname "$0")/.."
for spec in "$@"; do
    v="${spec%%@*}"; envs=""; [ "$spec" != "$v" ] && envs="${spec#*@}"
    unset CSI_PF_DIST
    for kv in ${envs//,/ }; do export "$kv"; done
    if [ "$v" = "base" ]; then unset CSI_B200_LIB; else export CSI_B200_LIB="$PWD/$v"; fi
    python bench.py --steps ${AB_STEPS:-3} --warmup 3 --no-e2e --no-cpu --no-configs ${AB_ARGS} > /tmp/ab.json 2> /tmp/ab.err || { echo "== $spec FAILED"; tail -5 /tmp/ab.err; continue; }
    python - "$spec" <<'EOF'
import json, sys
d = json.loads(open('/tmp/ab.json').read().strip().splitlines()[-1])
print(f"== {sys.argv[1]:28s} value={d['value']:.4e}  launch_ms={d['roofline']['launch_ms']:.4f}  frac={d['roofline']['frac']:.4f}  sm_mhz={d['clocks']['sm_mhz']}")
EOF
done
