#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 300 python bench.py --workload small --steps 2 --warmup 3 > gpurun_out/bench_small.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_c2.log 2>&1; echo "rc=$?" >> gpurun_out/bench_c2.log
tail -n 3 gpurun_out/smoke.log; tail -n 2 gpurun_out/bench_small.log; tail -n 3 gpurun_out/bench_c2.log
