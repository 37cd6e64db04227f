#!/bin/bash
mkdir -p gpurun_out
run() { # name lib envs...
  name=$1; lib=$2; shift 2
  env "$@" IRSDE_B200_LIB=$PWD/$lib IRSDE_PROFILE_DUMP=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ab.log 2> gpurun_out/prof_$name.txt
  python - "$name" <<'PY'
import json, sys
for line in open("gpurun_out/ab.log"):
    if line.startswith("{"):
        d=json.loads(line); print(sys.argv[1], "ms/chain",round(d["ms_per_step"],1), d["clocks"]["sm_mhz"], {k:round(v["ms_per_step"],3) for k,v in d["breakdown"].items()})
PY
}
run ref_b79 altlib/libirsde_b79c926.so A=1
run cur_fused image-restoration-sde_b200/libirsde_b200.so A=1
run cur_nofuse_out image-restoration-sde_b200/libirsde_b200.so IRSDE_LN_FUSE_OUT=0
run cur_nofuse image-restoration-sde_b200/libirsde_b200.so IRSDE_LN_FUSE=0
