#!/usr/bin/env bash
# session 2: re-capture of the T = 10 entries of profiles/traffic.json after the L2 prefetch went into K6-multi2
set -uo pipefail
mkdir -p gpurun_out
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum"
run() { local tag="$1"; local k="$2"; shift 2
  timeout 120 python profiles/prof_step.py "$@" > "gpurun_out/traffic_${tag}.plain.log" 2>&1 || { echo "plain run failed: $tag"; return; }
  timeout 300 ncu --metrics "$M" --clock-control none --cache-control none -k "regex:${k}" -s 40 -c 8 --csv \
    --log-file "gpurun_out/traffic_${tag}.csv" python profiles/prof_step.py "$@" > "gpurun_out/traffic_${tag}.ncu.log" 2>&1
  echo "$tag: $(tail -1 gpurun_out/traffic_${tag}.plain.log)"; }
for n in 2048 4096 8192 16384; do run step_kernel_T10_N$n step_ $n 10 80; done
