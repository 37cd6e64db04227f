#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fp32x3.py -q -m gpu -p no:cacheprovider -k "nf64" 2>&1 | tail -120 > gpurun_out/t_fp32x3_nf64.log
timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_refusion.py -q -m gpu -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/t_dropin_refusion.log
timeout 900 python -m pytest tests/test_gpu_bench_shapes.py -q -m gpu -p no:cacheprovider -k "denoising or qkv" 2>&1 | tail -60 > gpurun_out/t_attn.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/t_parity.log
tail -6 gpurun_out/t_dropin_refusion.log gpurun_out/t_attn.log gpurun_out/t_parity.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c2.log 2>&1; echo "rc=$?" >> gpurun_out/bench_c2.log
timeout 900 python bench.py --steps 2 --warmup 1 --precision fp32x3 --no-cpu > gpurun_out/bench_c2_fp32x3.log 2>&1; echo "rc=$?" >> gpurun_out/bench_c2_fp32x3.log
timeout 900 python bench.py --steps 1 --warmup 1 --workload c4 --no-cpu > gpurun_out/bench_c4.log 2>&1; echo "rc=$?" >> gpurun_out/bench_c4.log
for f in bench_c2 bench_c2_fp32x3 bench_c4; do echo "== $f"; tail -n 2 gpurun_out/$f.log | cut -c1-1500; done
