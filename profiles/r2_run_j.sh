#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
bash profiles/r2_traffic.sh
# experiment: 3 frame slots per env in the persistent kernel's stages -> 5 blocks per SM (aligned clips only: NOT a shippable build)
cd humanoid_b200/csrc && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -DPHC_PS_SLOTS=3 -DPHC_PS_BLOCKS=5 \
  -Xcompiler -fPIC,-fvisibility=hidden -shared -o ../libphc_b200.so phc_kernels.cu phc_host.cu phc_build.cu phc_peer.cu && cd ../..
timeout 600 python profiles/bench_persist.py 8192 16384 32768 65536 > gpurun_out/j_persist_slots3.log 2>&1; echo "persist slots3 rc=$?"; tail -6 gpurun_out/j_persist_slots3.log
