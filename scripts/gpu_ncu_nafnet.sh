#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/naf_launches.csv \
   python scripts/naf_steps_for_ncu.py 1 > gpurun_out/naf_ncu.log 2>&1
tail -n 2 gpurun_out/naf_ncu.log; wc -l gpurun_out/naf_launches.csv
