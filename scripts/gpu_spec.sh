#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bench_shapes.py tests/test_gpu_parity.py tests/test_gpu_refusion.py -q -m gpu -p no:cacheprovider -k "multi_tile or qkv or to_out or bf16 or tcgen05 or full_size or nafnet or latent" 2>&1 | tail -8 > gpurun_out/t_spec.log
tail -3 gpurun_out/t_spec.log | cut -c1-300
run() { # name lib specflag
  env IRSDE_B200_LIB=$PWD/$2 IRSDE_TC_EPI_SPEC=$3 IRSDE_PROFILE_DUMP=1 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/spec_$1.log 2> gpurun_out/spec_$1.err
  python - "$1" <<'PY'
import json, sys
for line in open("gpurun_out/spec_%s.log" % sys.argv[1]):
    if line.startswith("{"):
        d=json.loads(line); print(sys.argv[1], "ms/chain",round(d["ms_per_step"],1), d["clocks"]["sm_mhz"], "frac", round(d["roofline"]["frac"],3), {k:round(v["ms_per_step"],3) for k,v in d["breakdown"].items()})
PY
  grep "^PROF" gpurun_out/spec_$1.err > gpurun_out/prof_dump_spec_$1.txt
}
NEW=image-restoration-sde_b200/libirsde_b200.so
run s2 $NEW 2; run s3 $NEW 3; run s2b $NEW 2; run s3b $NEW 3
python scripts/prof_table.py gpurun_out/prof_dump_spec_s2.txt k3 | grep -E "32x32|64x64|total" | cut -c1-64,70-80 > gpurun_out/spec_gen_k1.txt
python scripts/prof_table.py gpurun_out/prof_dump_spec_s3.txt k3 | grep -E "32x32|64x64|total" | cut -c70-80 > gpurun_out/spec_spec_k1.txt
python scripts/prof_table.py gpurun_out/prof_dump_spec_s2b.txt k3 | grep -E "32x32|64x64|total" | cut -c70-80 > gpurun_out/spec_prev_k1.txt
paste -d'|' gpurun_out/spec_gen_k1.txt gpurun_out/spec_spec_k1.txt gpurun_out/spec_prev_k1.txt
