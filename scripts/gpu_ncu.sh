#!/bin/bash
# ncu evidence for the SHIPPED build (run under gpurun, 1 GPU):
#  (1) per-launch durations of >2 sampler steps of config 2, (2) --set full on one step's tcgen05 conv launches.
# Summaries only come back (the .ncu-rep stays in /tmp unless small).  scripts/ncu_summarize.py refuses a capture that
# holds no ROWS-mode launch (conv_tc_persist_kernel<*, 2>) - i.e. one that is not of the kernels bench.py times.
R=${ROUND:-r02}
mkdir -p gpurun_out
. image-restoration-sde_b200/BUILD_INFO 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 280 -c 300 --csv --log-file gpurun_out/${R}_launches_gpu_time.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu --no-graph --workload c2 > gpurun_out/ncu_list.log 2>&1
NK=${1:-82}
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_tc_persist -s 160 -c $NK -o /tmp/prof_tc -f \
   python bench.py --steps 1 --warmup 1 --no-cpu --no-graph --workload c2 > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/prof_tc.ncu-rep --page raw --csv > /tmp/prof_tc_raw.csv 2>/dev/null
python scripts/ncu_summarize.py /tmp/prof_tc_raw.csv --require 'conv_tc_persist_kernel<\d+, 2[,>]' --require 'conv_tc_persist_kernel<256, 0, 2' \
   --meta gpurun_out/${R}_conv_tc_ncu_full_one_step.meta.json commit=$commit csrc_sha256=$csrc_sha256 conv_tc_sha256=$conv_tc_sha256 \
   "command=ncu --set full --clock-control none -k regex:conv_tc_persist -s 160 -c $NK python bench.py --steps 1 --warmup 1 --no-cpu --no-graph --workload c2" \
   > gpurun_out/${R}_conv_tc_ncu_full_one_step.csv 2>gpurun_out/ncu_sum.err; echo "summarize rc=$?" >> gpurun_out/ncu_sum.err
# the other kernels of the step: HBM-bound LayerNorm / linear-attention k,v pass / sampler update
timeout 600 ncu --set full --clock-control none -k regex:"layernorm|la_kv|la_combine|la_fold|sde_update" -s 40 -c 40 -o /tmp/prof_hbm -f \
   python bench.py --steps 1 --warmup 1 --no-cpu --no-graph --workload c2 > gpurun_out/ncu_hbm.log 2>&1
ncu -i /tmp/prof_hbm.ncu-rep --page raw --csv > /tmp/prof_hbm_raw.csv 2>/dev/null
python scripts/ncu_summarize.py /tmp/prof_hbm_raw.csv --meta gpurun_out/${R}_hbm_kernels_ncu_full.meta.json commit=$commit csrc_sha256=$csrc_sha256 conv_tc_sha256=$conv_tc_sha256 \
   > gpurun_out/${R}_hbm_kernels_ncu_full.csv 2>>gpurun_out/ncu_sum.err
ls -la /tmp/*.ncu-rep gpurun_out/ | tail -n 16; cat gpurun_out/ncu_sum.err
