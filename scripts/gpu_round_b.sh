#!/bin/bash
# round-2 GPU pass B: fp32x3 mode tests, fixed benchmark-shape tests, drop-in test, bench lines (bf16, fp32x3, reference arms)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fp32x3.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -60 > gpurun_out/t_fp32x3.log
timeout 900 python -m pytest tests/test_gpu_bench_shapes.py tests/test_gpu_dropin.py -q -m gpu -p no:cacheprovider 2>&1 | tail -150 > gpurun_out/t_bench_shapes.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/t_parity.log
tail -4 gpurun_out/t_fp32x3.log gpurun_out/t_bench_shapes.log gpurun_out/t_parity.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c2.log 2>&1; echo "rc=$?" >> gpurun_out/bench_c2.log
timeout 900 python bench.py --steps 2 --warmup 1 --precision fp32x3 --no-cpu > gpurun_out/bench_c2_fp32x3.log 2>&1; echo "rc=$?" >> gpurun_out/bench_c2_fp32x3.log
timeout 600 python bench.py --impl reference --device cuda --steps 1 > gpurun_out/bench_ref_cuda.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref_cuda.log
timeout 600 python bench.py --impl reference --device cuda --no-tf32 --steps 1 > gpurun_out/bench_ref_cuda_notf32.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref_cuda_notf32.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_cpu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref_cpu.log
for f in bench_c2 bench_c2_fp32x3 bench_ref_cuda bench_ref_cuda_notf32 bench_ref_cpu; do echo "== $f"; tail -n 2 gpurun_out/$f.log | cut -c1-900; done
