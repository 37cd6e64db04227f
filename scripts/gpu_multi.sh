#!/bin/bash
# N-GPU weak-scaling bench exactly as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_multi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n$N.log
tail -n 3 gpurun_out/bench_n$N.log | cut -c1-900
