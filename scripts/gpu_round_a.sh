#!/bin/bash
# round-2 GPU pass: all GPU tests, smoke, default bench line, then the ncu captures of the shipped build
bash scripts/gpu_tests.sh
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c2.log 2>&1; echo "rc=$?" >> gpurun_out/bench_c2.log
tail -n 3 gpurun_out/smoke.log; tail -n 2 gpurun_out/bench_c2.log | cut -c1-1800
bash scripts/gpu_ncu.sh
