#!/bin/bash
# GPU test driver used under gpurun: separate processes so a trapped kernel cannot poison later groups.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
timeout 900 python -m pytest tests/test_gpu_bench_shapes.py -q -m gpu -p no:cacheprovider 2>&1 | tail -150 > gpurun_out/t_bench_shapes.log
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_bench_shapes.py -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/t_rest.log
tail -5 gpurun_out/t_bench_shapes.log gpurun_out/t_rest.log
