#!/bin/bash
# same-box ABAB of two libraries: altlib/libirsde_prev.so (previous kernels, git-ignored, shipped) vs the in-tree build
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bench_shapes.py tests/test_gpu_parity.py -q -m gpu -k "tcgen05 or qkv or to_out or nf64_forward_bf16 or unet_forward_bf16" -x -p no:cacheprovider 2>&1 | tail -3
run() { # name lib
  env IRSDE_B200_LIB=$PWD/$2 IRSDE_PROFILE_DUMP=1 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/ab_$1.log 2> gpurun_out/ab_$1.err
  python - "$1" <<'PY'
import json, sys
for line in open("gpurun_out/ab_%s.log" % sys.argv[1]):
    if line.startswith("{"):
        d=json.loads(line); print(sys.argv[1], "ms/chain",round(d["ms_per_step"],1), d["clocks"]["sm_mhz"], "frac", round(d["roofline"]["frac"],3), {k:round(v["ms_per_step"],3) for k,v in d["breakdown"].items()})
PY
  grep "^PROF" gpurun_out/ab_$1.err > gpurun_out/prof_dump_$1.txt
}
run prev altlib/libirsde_prev.so
run new image-restoration-sde_b200/libirsde_b200.so
run prev2 altlib/libirsde_prev.so
run new2 image-restoration-sde_b200/libirsde_b200.so
python scripts/prof_table.py gpurun_out/prof_dump_prev.txt to_qkv to_out res_conv | tail -30
python scripts/prof_table.py gpurun_out/prof_dump_new.txt to_qkv to_out res_conv | tail -30
