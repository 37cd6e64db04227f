#!/bin/bash
# A/B of env-switchable engine options on the bench breakdown
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -k "bf16 or unet_forward" -x -p no:cacheprovider 2>&1 | tail -5
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg IRSDE_PROFILE_DUMP=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ab.log 2> gpurun_out/prof_$(echo $cfg | tr -c 'A-Za-z0-9\n' '_').txt
  python - <<'PY'
import json
for line in open("gpurun_out/ab.log"):
    if line.startswith("{"):
        d=json.loads(line); print("value",round(d["value"],3),"ms/step",round(d["ms_per_step"],1), d["clocks"]["sm_mhz"], {k:round(v["ms_per_step"],3) for k,v in d["breakdown"].items()})
PY
done
