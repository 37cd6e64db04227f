#!/bin/bash
mkdir -p gpurun_out
for lib in "$@"; do
  echo "== $lib"
  IRSDE_B200_LIB=$PWD/$lib IRSDE_PROFILE_DUMP=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ab.log 2> gpurun_out/prof_$(basename $lib .so).txt
  python - <<'PY'
import json
for line in open("gpurun_out/ab.log"):
    if line.startswith("{"):
        d=json.loads(line); print("value",round(d["value"],3),"ms/step",round(d["ms_per_step"],1), d["clocks"]["sm_mhz"], {k:round(v["ms_per_step"],3) for k,v in d["breakdown"].items()})
PY
  python scripts/prof_table.py gpurun_out/prof_$(basename $lib .so).txt "to_qkv.weight 256" "ups.3.0.res" | head -3 2>/dev/null
done
