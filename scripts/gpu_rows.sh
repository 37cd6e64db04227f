#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -k "tcgen05 or bf16 or unet_forward" -x -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/t_quick.log
tail -n 4 gpurun_out/t_quick.log
bash scripts/gpu_micro.sh
