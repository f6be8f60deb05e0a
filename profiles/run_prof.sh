#!/usr/bin/env bash
# usage (on the GPU box, via gpurun): profiles/run_prof.sh <tag> [num_envs] [kernel-regex]
# plain run first (must exit 0), then ONE ncu --set full capture of the fused step kernel.
set -uo pipefail
tag="$1"; n="${2:-4096}"; k="${3:-step_fast}"
timeout 120 python profiles/prof_step.py $n > "gpurun_out/plain_${tag}.log" 2>&1 || { echo "plain run failed"; cat "gpurun_out/plain_${tag}.log"; exit 1; }
cat "gpurun_out/plain_${tag}.log"
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:${k}" -s 20 -c 1 \
  -o "gpurun_out/prof_${tag}" python profiles/prof_step.py $n > "gpurun_out/ncu_${tag}.log" 2>&1
tail -2 "gpurun_out/ncu_${tag}.log"
