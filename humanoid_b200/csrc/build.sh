#!/usr/bin/env bash
# Builds humanoid_b200/libphc_b200.so in-tree for sm_100a (cross-compiles without a GPU).
#   -fmad=false : a*b+c stays two roundings, as in ATen's op-by-op evaluation (phc_math.cuh)
#   -lineinfo   : ncu source page maps to these files
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${here}/../libphc_b200.so"
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
  -Xcompiler -fPIC,-fvisibility=hidden -shared ${PHC_NVCC_EXTRA:-} \
  -o "${out}" "${here}/phc_kernels.cu" "${here}/phc_host.cu" "${here}/phc_build.cu" "${here}/phc_peer.cu"
echo "built ${out}"
