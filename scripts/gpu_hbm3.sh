#!/bin/bash
# Last check of the round: IRSDE_HBM_NEW=31 (adds the lean k/v pass and the 8-warp merge) against the suite-validated
# setting 7, then a driver-style line and the bf16 GPU tests under the better passing setting.  Every step writes its
# output as it finishes (the call may be cut by the GPU budget).
mkdir -p gpurun_out
export HBM_PROBE_ATTN_ONLY=1 HBM_PROBE_TOL_BF16=1e-2
probe() { env IRSDE_HBM_NEW=$1 timeout 200 python scripts/hbm_probe.py run /tmp/probe_$1.pt 2>&1 | tail -1; }
probe 7 > gpurun_out/hbm3_probe.txt; probe 31 >> gpurun_out/hbm3_probe.txt
GOOD=7
if python scripts/hbm_probe.py cmp /tmp/probe_7.pt /tmp/probe_31.pt >> gpurun_out/hbm3_probe.txt 2>&1; then GOOD=31; else
  probe 23 >> gpurun_out/hbm3_probe.txt
  python scripts/hbm_probe.py cmp /tmp/probe_7.pt /tmp/probe_23.pt >> gpurun_out/hbm3_probe.txt 2>&1 && GOOD=23
fi
echo "GOOD=$GOOD" >> gpurun_out/hbm3_probe.txt; cat gpurun_out/hbm3_probe.txt
export IRSDE_HBM_NEW=$GOOD
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/hbm3_bench_c2.log 2> gpurun_out/hbm3_bench_c2.err
grep '^{' gpurun_out/hbm3_bench_c2.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c2', d['build'], round(d['value'],3), 'img/s', round(d['ms_per_step'],1), 'ms', {k: round(v['ms_per_step'],3) for k,v in d['breakdown'].items()})"
timeout 400 python -m pytest tests/test_gpu_bench_shapes.py tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -k "bf16 or full_size or denoising or sharded" 2>&1 | tail -4 | tee gpurun_out/hbm3_tests.log
env IRSDE_HBM_NEW=7 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/hbm3_bench_c2_mask7.log 2>/dev/null
grep '^{' gpurun_out/hbm3_bench_c2_mask7.log | tail -1 | cut -c1-160
