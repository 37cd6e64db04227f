#!/bin/bash
# quick GPU check: tensor-core + bf16 tests, then the bench (no CPU baseline)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -k "tcgen05 or bf16 or chain_fp32" -x -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/t_quick.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_c2.log 2>&1; echo "rc=$?" >> gpurun_out/bench_c2.log
if [ -n "$1" ]; then IRSDE_TC_PERSIST=0 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_c2_v1.log 2>&1; fi
tail -n 5 gpurun_out/t_quick.log; python - <<'PY'
import json
for f in ("gpurun_out/bench_c2.log","gpurun_out/bench_c2_v1.log"):
    try:
        for line in open(f):
            if line.startswith("{"):
                d=json.loads(line); print(f, "value",round(d["value"],3),"e2e",round(d["e2e"]["value"],3),"ms/step",round(d["ms_per_step"],1), "roof",d["roofline"] and round(d["roofline"]["achieved"],1), d["clocks"])
                print({k:(round(v["ms_per_step"],3), v["tflops"] and round(v["tflops"],1)) for k,v in d["breakdown"].items()})
    except Exception as e: print(f, e)
PY
