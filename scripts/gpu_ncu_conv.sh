#!/bin/bash
# the --set full conv capture of scripts/gpu_ncu.sh alone (the launch list and the HBM-kernel capture already came back)
R=${ROUND:-r02}
mkdir -p gpurun_out
. image-restoration-sde_b200/BUILD_INFO 2>/dev/null
NK=${1:-82}
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_tc_persist -s 160 -c $NK -o /tmp/prof_tc -f \
   python bench.py --steps 1 --warmup 1 --no-cpu --no-graph --workload c2 > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/prof_tc.ncu-rep --page raw --csv > /tmp/prof_tc_raw.csv 2>/dev/null
python scripts/ncu_summarize.py /tmp/prof_tc_raw.csv --require 'conv_tc_persist_kernel<\d+, 2[,>]' --require 'conv_tc_persist_kernel<256, 0, 2' \
   --meta gpurun_out/${R}_conv_tc_ncu_full_one_step.meta.json commit=$commit csrc_sha256=$csrc_sha256 conv_tc_sha256=$conv_tc_sha256 \
   "command=ncu --set full --clock-control none -k regex:conv_tc_persist -s 160 -c $NK python bench.py --steps 1 --warmup 1 --no-cpu --no-graph --workload c2" \
   > gpurun_out/${R}_conv_tc_ncu_full_one_step.csv 2>gpurun_out/ncu_sum.err; echo "summarize rc=$?" >> gpurun_out/ncu_sum.err
cat gpurun_out/ncu_sum.err; wc -l gpurun_out/${R}_conv_tc_ncu_full_one_step.csv
