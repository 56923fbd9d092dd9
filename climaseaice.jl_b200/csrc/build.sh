#!/bin/sh
# Builds libclimaseaice_b200.so in-tree for sm_100a.  -fmad=false: the reference's Float64
# operation sequence must be reproduced exactly; FMA appears only where written explicitly.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=${CSI_OUT:-../libclimaseaice_b200.so}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false \
      -Xcompiler -fPIC -Xcompiler -O2 -Xcompiler -ffp-contract=off -shared -DCSI_FUSED_MINB=${CSI_FUSED_MINB:-3} \
      ${CSI_NVCC_EXTRA} \
      -o $OUT csi_api.cu csi_unfused.cu csi_halo.cu csi_advection.cu csi_reduce.cu csi_thermo.cu csi_fused.cu \
      -cudart static -ldl
echo "built $OUT"
