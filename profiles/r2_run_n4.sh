#!/usr/bin/env bash
# 4 GPUs, final binary of round 2: the bench as the driver launches it
set -uo pipefail
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29631 \
  bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/n4_bench.json 2> gpurun_out/n4_bench.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/n4_bench.json"))
e=d["e2e"]; print("value %.3g"%d["value"], "us/step %.2f"%(d["ms_per_step"]*1e3), "e2e %.3g"%e["value"], "e2e us %.0f"%e["us_per_step"], e["output_path"][:8], "chunks", e["chunks"], "floor us %.0f"%e["floor"]["us_per_step"], "e2e/floor %.2f"%e["floor"]["e2e_over_floor"])
print({k:v for k,v in d["rms"].items() if k.startswith(("peer","ranks","exchange"))})
for k,v in d.get("configs",{}).items(): print(k, "%.2f us"%v["us_per_step"], "frac %.3f"%v["frac"], ("rollout %.2f us/step"%v["rollout"]["us_per_step"]) if "rollout" in v else "")
PY
