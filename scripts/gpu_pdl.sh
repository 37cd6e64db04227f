#!/bin/bash
# programmatic dependent launch: correctness (whole GPU suite with PDL on) + same-box A/B via IRSDE_PDL
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -4
run() { # name pdl
  env IRSDE_PDL=$2 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ab_$1.log 2>&1
  python - "$1" <<'PY'
import json, sys
ok = False
for line in open("gpurun_out/ab_%s.log" % sys.argv[1]):
    if line.startswith("{"):
        ok = True
        d=json.loads(line); print(sys.argv[1], "ms/chain",round(d["ms_per_step"],1), d["clocks"]["sm_mhz"], "frac", round(d["roofline"]["frac"],3), {k:round(v["ms_per_step"],3) for k,v in d["breakdown"].items()})
if not ok: print(sys.argv[1], "FAILED", open("gpurun_out/ab_%s.log" % sys.argv[1]).read()[-600:])
PY
}
run pdl0 0
run pdl1 1
run pdl0b 0
run pdl1b 1
for p in 0 1; do echo "naf pdl=$p"; IRSDE_PDL=$p timeout 300 python scripts/time_latent.py 1 1024 2>&1 | tail -1 | cut -c1-330; done
