#!/usr/bin/env bash
# 8 GPUs: the bench as the driver launches it (configs 3 / 5 strong-scaled, peer check, e2e + floor with 8 ranks), once with the
# output path chosen by the library and once with each path pinned
set -uo pipefail
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; numactl -H 2>/dev/null | head -5
run() { # tag, extra env
  local tag="$1"; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 \
    bench.py --gpus 8 --steps 20 --warmup 5 $EXTRA > gpurun_out/n8_${tag}.json 2> gpurun_out/n8_${tag}.err; echo "$tag rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/n8_${tag}.json"))
e=d["e2e"]; print("${tag}", "value %.3g"%d["value"], "us/step %.2f"%(d["ms_per_step"]*1e3), "e2e %.3g"%e["value"], "e2e us %.0f"%e["us_per_step"], e["output_path"][:8], "floor us %.0f"%e["floor"]["us_per_step"], "d2h-only %.0f"%e["floor"]["d2h_only_us"], "h2d-only %.0f"%e["floor"]["h2d_only_us"], "e2e/floor %.2f"%e["floor"]["e2e_over_floor"])
print({k:v for k,v in d["rms"].items() if k.startswith(("peer","ranks","exchange"))})
for k,v in d.get("configs",{}).items(): print(k, "%.2f us"%v["us_per_step"], "frac %.3f"%v["frac"], ("rollout %.2f us/step"%v["rollout"]["us_per_step"]) if "rollout" in v else "")
PY
}
EXTRA="" run auto PHC_X=1
EXTRA="--no-extra-configs" run direct PHC_HOST_PATH=direct
EXTRA="--no-extra-configs" run staged PHC_HOST_PATH=staged
