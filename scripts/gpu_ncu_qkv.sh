#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_persist -s 2 -c 1 -o /tmp/prof_qkv -f python scripts/ncu_qkv.py 1 > gpurun_out/ncu_qkv.log 2>&1
ncu -i /tmp/prof_qkv.ncu-rep --page source --csv > gpurun_out/src_qkv.csv 2>/dev/null
ncu -i /tmp/prof_qkv.ncu-rep --page details --csv > gpurun_out/details_qkv.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_persist -s 2 -c 1 -o /tmp/prof_qkv0 -f python scripts/ncu_qkv.py 0 > gpurun_out/ncu_qkv0.log 2>&1
ncu -i /tmp/prof_qkv0.ncu-rep --page details --csv > gpurun_out/details_qkv0.csv 2>/dev/null
ncu -i /tmp/prof_qkv0.ncu-rep --page source --csv > gpurun_out/src_qkv0.csv 2>/dev/null
ls -la gpurun_out/*qkv*; tail -3 gpurun_out/ncu_qkv.log
