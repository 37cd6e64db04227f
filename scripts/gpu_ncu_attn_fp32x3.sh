#!/bin/bash
R=${ROUND:-r02}
mkdir -p gpurun_out
. image-restoration-sde_b200/BUILD_INFO 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:fullattn -s 1 -c 2 -o /tmp/prof_attn -f python scripts/ncu_attn_fp32x3.py attn > gpurun_out/ncu_attn.log 2>&1
ncu -i /tmp/prof_attn.ncu-rep --page raw --csv > /tmp/prof_attn_raw.csv 2>/dev/null
python scripts/ncu_summarize.py /tmp/prof_attn_raw.csv --require fullattn_mma_kernel --meta gpurun_out/${R}_fullattn_ncu_full.meta.json commit=$commit csrc_sha256=$csrc_sha256 conv_tc_sha256=$conv_tc_sha256 > gpurun_out/${R}_fullattn_ncu_full.csv 2> gpurun_out/ncu_attn.err
ncu -i /tmp/prof_attn.ncu-rep --page details --csv 2>/dev/null | grep -E "Pipe|pipe|Duration|Issue Slots|Executed Ipc" | head -40 > gpurun_out/${R}_fullattn_details.txt
timeout 900 ncu --set full --clock-control none -k regex:conv_tc_persist -s 74 -c 74 -o /tmp/prof_f3 -f python scripts/ncu_attn_fp32x3.py f3 > gpurun_out/ncu_f3.log 2>&1
ncu -i /tmp/prof_f3.ncu-rep --page raw --csv > /tmp/prof_f3_raw.csv 2>/dev/null
python scripts/ncu_summarize.py /tmp/prof_f3_raw.csv --require 'conv_tc_persist_kernel<\d+, 3[,>]' --meta gpurun_out/${R}_fp32x3_conv_ncu_full_one_forward.meta.json commit=$commit csrc_sha256=$csrc_sha256 conv_tc_sha256=$conv_tc_sha256 > gpurun_out/${R}_fp32x3_conv_ncu_full_one_forward.csv 2> gpurun_out/ncu_f3.err
cat gpurun_out/ncu_attn.err gpurun_out/ncu_f3.err; wc -l gpurun_out/${R}_f*.csv; cat gpurun_out/${R}_fullattn_details.txt | cut -c1-200 | head -30
timeout 900 python bench.py --steps 2 --warmup 1 --workload c3 --no-cpu > gpurun_out/bench_c3_n1.log 2>&1
timeout 900 python bench.py --steps 1 --warmup 1 --workload c5 --no-cpu > gpurun_out/bench_c5_n1.log 2>&1
for f in bench_c3_n1 bench_c5_n1; do grep '^{' gpurun_out/$f.log | tail -1 | cut -c1-330; done
