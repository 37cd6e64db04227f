#!/bin/bash
# ncu --set full on the HBM-bound kernels of one sampler step (LayerNorm, linear-attention k/v pass, sampler update)
mkdir -p gpurun_out
timeout 420 ncu --set full --clock-control none -k regex:"layernorm_vec|la_kv_mma|sde_update" -c 30 -o /tmp/prof_hbm -f \
   python bench.py --steps 1 --warmup 1 --no-cpu --no-graph --workload c2 > gpurun_out/ncu_hbm.log 2>&1
ncu -i /tmp/prof_hbm.ncu-rep --page raw --csv > /tmp/prof_hbm_raw.csv 2>/dev/null
python scripts/ncu_summarize.py /tmp/prof_hbm_raw.csv > gpurun_out/prof_hbm_summary.csv 2>gpurun_out/ncu_hbm.err
wc -l gpurun_out/prof_hbm_summary.csv; head -c 600 gpurun_out/prof_hbm_summary.csv
