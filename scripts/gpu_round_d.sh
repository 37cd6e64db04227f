#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fp32x3.py -q -m gpu -p no:cacheprovider 2>&1 | tail -80 > gpurun_out/t_fp32x3.log
tail -5 gpurun_out/t_fp32x3.log
timeout 900 python bench.py --steps 2 --warmup 1 --precision fp32x3 --no-cpu > gpurun_out/bench_c2_fp32x3.log 2>&1; echo "rc=$?" >> gpurun_out/bench_c2_fp32x3.log
timeout 900 python bench.py --steps 2 --warmup 1 --workload c3 --no-cpu > gpurun_out/bench_c3_n1.log 2>&1; echo "rc=$?" >> gpurun_out/bench_c3_n1.log
timeout 900 python bench.py --steps 1 --warmup 1 --workload c5 --no-cpu > gpurun_out/bench_c5_n1.log 2>&1; echo "rc=$?" >> gpurun_out/bench_c5_n1.log
for f in bench_c2_fp32x3 bench_c3_n1 bench_c5_n1; do echo "== $f"; tail -n 2 gpurun_out/$f.log | cut -c1-1200; done
