#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_ab.sh 2>&1 | tail -80 > gpurun_out/ab_summary.txt
cat gpurun_out/ab_summary.txt | cut -c1-260
timeout 900 python -m pytest tests/test_gpu_fp32x3.py -q -m gpu -p no:cacheprovider -k "nf64 or 1536 or chain" 2>&1 | tail -70 > gpurun_out/t_fp32x3.log
grep -n "forward fp32x3\|raw random\|assisted\|passed\|failed\|final_res_block" gpurun_out/t_fp32x3.log | cut -c1-200
timeout 900 python bench.py --steps 2 --warmup 1 --precision fp32x3 --no-cpu > gpurun_out/bench_c2_fp32x3.log 2>&1; echo "rc=$?" >> gpurun_out/bench_c2_fp32x3.log
tail -n 2 gpurun_out/bench_c2_fp32x3.log | cut -c1-300
