#!/bin/bash
# CTA-pair (cta_group::2) variant: correctness first (separate process: a trap poisons the context), then ABAB pair off/on
mkdir -p gpurun_out
IRSDE_TC_PAIR=1 timeout 300 python -m pytest tests/test_gpu_bench_shapes.py -q -m gpu -p no:cacheprovider -x -k "1536 or mid or 768" 2>&1 | tail -25 > gpurun_out/t_pair_convs.log
tail -6 gpurun_out/t_pair_convs.log | cut -c1-300
if grep -q "passed" gpurun_out/t_pair_convs.log && ! grep -q "failed\|error" gpurun_out/t_pair_convs.log; then
  IRSDE_TC_PAIR=1 timeout 600 python -m pytest tests/test_gpu_bench_shapes.py tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -k "bf16 or tcgen05 or full_size" 2>&1 | tail -8 > gpurun_out/t_pair_nets.log
  tail -3 gpurun_out/t_pair_nets.log | cut -c1-300
  run() { # name pairflag
    env IRSDE_TC_PAIR=$2 IRSDE_PROFILE_DUMP=1 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/pair_$1.log 2> gpurun_out/pair_$1.err
    python - "$1" <<'PY'
import json, sys
for line in open("gpurun_out/pair_%s.log" % sys.argv[1]):
    if line.startswith("{"):
        d=json.loads(line); print(sys.argv[1], "ms/chain",round(d["ms_per_step"],1), d["clocks"]["sm_mhz"], "frac", round(d["roofline"]["frac"],3), {k:round(v["ms_per_step"],3) for k,v in d["breakdown"].items()})
PY
    grep "^PROF" gpurun_out/pair_$1.err > gpurun_out/prof_dump_pair_$1.txt
  }
  run off 0; run on 1; run off2 0; run on2 1
fi
