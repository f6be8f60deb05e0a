#!/usr/bin/env bash
# Round 2: steady-state DRAM traffic of the step kernel per config.
# One-pass ncu capture (no kernel replay, --cache-control none: caches are left as the ring of buffer sets leaves
# them), launches 40..47 of a 60-launch run over a 17-set ring, i.e. after two laps.  profiles/traffic_from_ncu.py
# turns the CSVs into profiles/traffic.json (mean bytes per launch) and profiles/r2_traffic.md.
set -uo pipefail
mkdir -p gpurun_out
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum"
run() { # tag regex args...
  local tag="$1"; local k="$2"; shift 2
  timeout 120 python profiles/prof_step.py "$@" > "gpurun_out/traffic_${tag}.plain.log" 2>&1 || { echo "plain run failed: $tag"; cat "gpurun_out/traffic_${tag}.plain.log"; return; }
  timeout 300 ncu --metrics "$M" --clock-control none --cache-control none -k "regex:${k}" -s 40 -c 8 --csv \
    --log-file "gpurun_out/traffic_${tag}.csv" python profiles/prof_step.py "$@" > "gpurun_out/traffic_${tag}.ncu.log" 2>&1
  echo "$tag: $(cat gpurun_out/traffic_${tag}.plain.log | tail -1)"
}
# launches 40..47 of an 80-launch run; every config bench.py reports at 1 / 2 / 4 / 8 GPUs
for n in 4096 8192 16384 32768 65536; do run step_kernel_T1_N$n step_ $n 1 80; done
run config4_T1_N4096 step_ 4096 1 80 config4
for n in 2048 4096 8192 16384; do run step_kernel_T10_N$n step_ $n 10 80; done
python profiles/traffic_from_ncu.py gpurun_out > gpurun_out/traffic_summary.md; cat gpurun_out/traffic_summary.md
