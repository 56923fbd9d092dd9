#!/bin/sh
# Builds libclimaseaice_b200.so in-tree for sm_100a.  -fmad=false: the reference's Float64
# operation sequence must be reproduced exactly; FMA appears only where written explicitly.
# The translation units compile in parallel (the fused kernel's instantiations dominate).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=${CSI_OUT:-../libclimaseaice_b200.so}
OBJ=${CSI_OBJ:-/tmp/csi_b200_obj$(echo "$OUT" | tr -c "A-Za-z0-9" "_")}
mkdir -p "$OBJ"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -Xcompiler -O2 -Xcompiler -ffp-contract=off -DCSI_FUSED_MINB=${CSI_FUSED_MINB:-3} ${CSI_NVCC_EXTRA}"
pids=""
for f in csi_api csi_unfused csi_halo csi_advection csi_reduce csi_thermo csi_bench csi_fused; do
    $NVCC $FLAGS -c -o "$OBJ/$f.o" $f.cu &
    pids="$pids $!"
done
for p in $pids; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT "$OBJ"/csi_api.o "$OBJ"/csi_unfused.o "$OBJ"/csi_halo.o "$OBJ"/csi_advection.o \
      "$OBJ"/csi_reduce.o "$OBJ"/csi_thermo.o "$OBJ"/csi_bench.o "$OBJ"/csi_fused.o -cudart static -ldl
echo "built $OUT"
