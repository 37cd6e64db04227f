#!/bin/bash
# final evidence pass of the round on the shipped build: full GPU suite, smoke, driver-style bench lines, ncu captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1; nproc >> gpurun_out/smi.txt
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/t_all.log; tail -4 gpurun_out/t_all.log | cut -c1-300
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2_n1.log 2>&1; echo "rc=$?" >> gpurun_out/bench_c2_n1.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_cpu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref_cpu.log
timeout 900 python bench.py --steps 3 --warmup 2 --precision fp32x3 --no-cpu > gpurun_out/bench_c2_fp32x3.log 2>&1; echo "rc=$?" >> gpurun_out/bench_c2_fp32x3.log
timeout 900 python bench.py --steps 2 --warmup 1 --precision fp32 --no-cpu > gpurun_out/bench_c2_fp32.log 2>&1; echo "rc=$?" >> gpurun_out/bench_c2_fp32.log
timeout 900 python bench.py --steps 2 --warmup 1 --workload c4 --no-cpu > gpurun_out/bench_c4_n1.log 2>&1; echo "rc=$?" >> gpurun_out/bench_c4_n1.log
for f in bench_c2_n1 bench_ref_cpu bench_c2_fp32x3 bench_c2_fp32 bench_c4_n1; do echo "== $f"; grep '^{' gpurun_out/$f.log | tail -1 | cut -c1-420; done
bash scripts/gpu_ncu.sh
