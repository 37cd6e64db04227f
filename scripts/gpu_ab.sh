#!/bin/bash
# same-box A/B of two libraries: altlib/libirsde_base.so (a copy of the previous build, git-ignored) vs the in-tree build (bench breakdown + a few bf16 tests)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -k "bf16 or tcgen05 or unet_forward or nafnet or chain" -x -p no:cacheprovider 2>&1 | tail -3
run() { # name lib
  env IRSDE_B200_LIB=$PWD/$2 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ab_$1.log 2>&1
  python - "$1" <<'PY'
import json, sys
for line in open("gpurun_out/ab_%s.log" % sys.argv[1]):
    if line.startswith("{"):
        d=json.loads(line); print(sys.argv[1], "ms/chain",round(d["ms_per_step"],1), d["clocks"]["sm_mhz"], "frac", round(d["roofline"]["frac"],3), {k:round(v["ms_per_step"],3) for k,v in d["breakdown"].items()})
PY
}
run base altlib/libirsde_base.so
run new image-restoration-sde_b200/libirsde_b200.so
run base2 altlib/libirsde_base.so
run new2 image-restoration-sde_b200/libirsde_b200.so
