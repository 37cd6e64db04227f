#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bench_shapes.py tests/test_gpu_parity.py tests/test_gpu_refusion.py -q -m gpu -p no:cacheprovider -k "bf16 or full_size or nafnet or latent or sharded" 2>&1 | tail -5 > gpurun_out/t_quick.log
tail -3 gpurun_out/t_quick.log | cut -c1-300
for i in 1 2; do
env IRSDE_PROFILE_DUMP=1 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/quick_$i.log 2> gpurun_out/quick_$i.err
python - "$i" <<'PY'
import json, sys
for line in open("gpurun_out/quick_%s.log" % sys.argv[1]):
    if line.startswith("{"):
        d=json.loads(line); print("run", sys.argv[1], "ms/chain",round(d["ms_per_step"],1), d["clocks"]["sm_mhz"], "frac", round(d["roofline"]["frac"],3), {k:round(v["ms_per_step"],3) for k,v in d["breakdown"].items()})
PY
done
grep "^PROF" gpurun_out/quick_1.err > gpurun_out/prof_dump_quick.txt
python scripts/prof_table.py gpurun_out/prof_dump_quick.txt linattn fold norm | tail -40
