#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_bench_shapes.py -q -m gpu -p no:cacheprovider -x -k "qkv or to_out or res_conv or multi_tile" 2>&1 | tail -25 > gpurun_out/t_wres_convs.log
tail -6 gpurun_out/t_wres_convs.log | cut -c1-300
if grep -q "passed" gpurun_out/t_wres_convs.log && ! grep -q "failed\|error" gpurun_out/t_wres_convs.log; then
  timeout 600 python -m pytest tests/test_gpu_bench_shapes.py tests/test_gpu_parity.py tests/test_gpu_refusion.py -q -m gpu -p no:cacheprovider -k "bf16 or tcgen05 or full_size or nafnet or latent" 2>&1 | tail -8 > gpurun_out/t_wres_nets.log
  tail -3 gpurun_out/t_wres_nets.log | cut -c1-300
  run() { # name flag
    env IRSDE_TC_WRES=$2 IRSDE_PROFILE_DUMP=1 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/wres_$1.log 2> gpurun_out/wres_$1.err
    python - "$1" <<'PY'
import json, sys
for line in open("gpurun_out/wres_%s.log" % sys.argv[1]):
    if line.startswith("{"):
        d=json.loads(line); print(sys.argv[1], "ms/chain",round(d["ms_per_step"],1), d["clocks"]["sm_mhz"], "frac", round(d["roofline"]["frac"],3), {k:round(v["ms_per_step"],3) for k,v in d["breakdown"].items()})
PY
    grep "^PROF" gpurun_out/wres_$1.err > gpurun_out/prof_dump_wres_$1.txt
  }
  run off 0; run on 1; run off2 0; run on2 1
  python scripts/prof_table.py gpurun_out/prof_dump_wres_off.txt k1 | tail -40 > gpurun_out/wres_off_k1.txt
  python scripts/prof_table.py gpurun_out/prof_dump_wres_on.txt k1 | tail -40 > gpurun_out/wres_on_k1.txt
  paste -d'|' <(cut -c1-64,70-80 gpurun_out/wres_off_k1.txt) <(cut -c70-80 gpurun_out/wres_on_k1.txt)
fi
