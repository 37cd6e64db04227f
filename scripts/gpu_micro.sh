#!/bin/bash
mkdir -p gpurun_out
IRSDE_TC_DEBUG=1 timeout 300 python scripts/micro_conv.py > gpurun_out/micro.log 2>&1
IRSDE_TC_DEBUG=1 IRSDE_TC_ROWS=0 timeout 300 python scripts/micro_conv.py > gpurun_out/micro_nopatch.log 2>&1
grep -E "case|TCDBG" gpurun_out/micro.log | cut -c1-330; echo ---- ; grep -E "TCDBG" gpurun_out/micro_nopatch.log | cut -c1-330
