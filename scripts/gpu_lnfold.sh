#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_gpu_bench_shapes.py tests/test_gpu_parity.py tests/test_gpu_refusion.py -q -m gpu -p no:cacheprovider -k "bf16 or full_size or latent or sharded or restorer or philox" 2>&1 | tail -100 > gpurun_out/t_lnfold.log
grep -E "passed|failed|rel-rms|output \(eps" gpurun_out/t_lnfold.log | grep -E "passed|failed|to_out.0.weight|norm\+res|output" | head -30 | cut -c1-200
run() { # name flag
  env IRSDE_LNFOLD=$2 IRSDE_PROFILE_DUMP=1 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/lnf_$1.log 2> gpurun_out/lnf_$1.err
  python - "$1" <<'PY'
import json, sys
for line in open("gpurun_out/lnf_%s.log" % sys.argv[1]):
    if line.startswith("{"):
        d=json.loads(line); print(sys.argv[1], "ms/chain",round(d["ms_per_step"],1), d["clocks"]["sm_mhz"], "frac", round(d["roofline"]["frac"],3), {k:round(v["ms_per_step"],3) for k,v in d["breakdown"].items()})
PY
  grep "^PROF" gpurun_out/lnf_$1.err > gpurun_out/prof_dump_lnf_$1.txt
}
run off 0; run on 1; run off2 0; run on2 1
python scripts/prof_table.py gpurun_out/prof_dump_lnf_on.txt norm to_qkv | grep -E "256x256|128x128|downs.0.2|ups.3.2|total" | cut -c1-110
