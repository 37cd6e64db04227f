#!/bin/bash
# multi-GPU pass (gpurun --gpus N): scripts/gpu_multi.sh N "wl:steps:warmup ..." [tests]
N=$1; shift
WLS=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/multi_smi_n$N.txt
if [ "$1" == "tests" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/t_multi.log; tail -3 gpurun_out/t_multi.log
fi
P=29610
for spec in $WLS; do
  IFS=: read wl steps warm <<< "$spec"
  P=$((P+1))
  if [ "$N" == "1" ]; then
    timeout 900 python bench.py --gpus 1 --steps $steps --warmup $warm --workload $wl --no-cpu > gpurun_out/bench_${wl}_n$N.log 2>&1
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps $steps --warmup $warm --workload $wl --no-cpu > gpurun_out/bench_${wl}_n$N.log 2>&1
  fi
  echo "== $wl N=$N rc=$?"; grep '^{' gpurun_out/bench_${wl}_n$N.log | tail -1 | cut -c1-400
done
