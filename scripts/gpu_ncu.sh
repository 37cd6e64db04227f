#!/bin/bash
# ncu evidence: (1) per-launch durations for >2 sampler steps, (2) --set full on one step's tcgen05 convs.
# Only text summaries come back (the .ncu-rep is deleted if it would blow the 64 MiB gpurun_out limit).
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 280 -c 300 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu --no-graph --workload c2 > gpurun_out/ncu_list.log 2>&1
NK=${1:-75}
timeout 1200 ncu --set full --clock-control none -k regex:conv_tc_persist -s 160 -c $NK -o /tmp/prof_tc -f \
   python bench.py --steps 1 --warmup 1 --no-cpu --no-graph --workload c2 > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/prof_tc.ncu-rep --page raw --csv > /tmp/prof_tc_raw.csv 2>/dev/null
python scripts/ncu_summarize.py /tmp/prof_tc_raw.csv > gpurun_out/prof_tc_summary.csv 2>gpurun_out/ncu_sum.err
ls -la /tmp/prof_tc.ncu-rep gpurun_out/ | tail -n 12
