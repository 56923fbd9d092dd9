#!/bin/bash
# A/B of library variants on the masked coastline case (the GEN instantiation): tools/ab_gen.sh base variants/x.so ...
cd "$(dirname "$0")/.."
for v in "$@"; do
    if [ "$v" = "base" ]; then unset CSI_B200_LIB; else export CSI_B200_LIB="$PWD/$v"; fi
    echo "== $v"
    python tools/profile_case.py 2048 150 fused coastline 2>&1 | grep "cell-updates"
    python tools/profile_case.py 2048 150 fused latlon 2>&1 | grep "cell-updates"
done
