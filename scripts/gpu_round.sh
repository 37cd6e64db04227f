#!/bin/bash
# one GPU round: parity tests, smoke, bench (config 2), ncu launch list + one full capture of the top kernel
mkdir -p gpurun_out
cat /sys/fs/cgroup/cpu.max > gpurun_out/cpuinfo.txt 2>&1; nproc >> gpurun_out/cpuinfo.txt; lscpu | head -20 >> gpurun_out/cpuinfo.txt
timeout 900 python -m pytest tests -q -m gpu -k "not tcgen05 and not bf16" -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/t_fp32.log
timeout 600 python -m pytest tests -q -m gpu -k "tcgen05" -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/t_tc.log
timeout 600 python -m pytest tests -q -m gpu -k "bf16" -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/t_bf16.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c2.log 2>&1; echo "rc=$?" >> gpurun_out/bench_c2.log
if [ "$1" == "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 330 --csv --log-file gpurun_out/launches.csv \
     python bench.py --steps 1 --warmup 1 --no-cpu --no-graph --workload c2 > gpurun_out/ncu_list.log 2>&1
  timeout 1200 ncu --set full --clock-control none -k regex:conv_tc -s 150 -c 75 -o gpurun_out/prof_tc -f \
     python bench.py --steps 1 --warmup 1 --no-cpu --no-graph --workload c2 > gpurun_out/ncu_full.log 2>&1
  ncu -i gpurun_out/prof_tc.ncu-rep --page raw --csv > gpurun_out/prof_tc_raw.csv 2>/dev/null
  ls -la gpurun_out/
fi
for f in t_fp32 t_tc t_bf16 smoke bench_c2; do echo "== $f"; tail -n 4 gpurun_out/$f.log | cut -c1-1500; done
