#!/bin/bash
# GPU test driver used under gpurun: separate processes so a trapped kernel cannot poison later groups.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu -k "not tcgen05 and not bf16" -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/t_fp32.log
timeout 600 python -m pytest tests -q -m gpu -k "tcgen05" -x -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/t_tc.log
timeout 600 python -m pytest tests -q -m gpu -k "bf16" -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/t_bf16.log
tail -5 gpurun_out/t_fp32.log gpurun_out/t_tc.log gpurun_out/t_bf16.log
