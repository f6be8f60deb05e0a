#!/usr/bin/env bash
# round 2: ncu launch list of the driver's bench command (headline config only), after the plain run has exited 0
set -uo pipefail
mkdir -p gpurun_out
B="--gpus 1 --steps 20 --warmup 5 --no-extra-configs --no-cpu-baseline --no-torch-cuda-baseline --no-long-run --e2e-steps 2 --repeats 5"
timeout 300 python bench.py $B > gpurun_out/l_plain.json 2> gpurun_out/l_plain.err; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_bench_launches_raw.csv \
  python bench.py $B > gpurun_out/l_ncu.json 2> gpurun_out/l_ncu.err; echo "ncu rc=$?"
wc -l gpurun_out/r2_bench_launches_raw.csv
python profiles/launch_summary.py gpurun_out/r2_bench_launches_raw.csv gpurun_out/r2_bench_launches.md
cat gpurun_out/r2_bench_launches.md | head -24
