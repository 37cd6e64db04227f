#!/bin/bash
mkdir -p gpurun_out
IRSDE_PROFILE_DUMP=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/bench_prof.log 2> gpurun_out/prof_dump.txt
grep -c PROF gpurun_out/prof_dump.txt
if [ -n "$1" ]; then
IRSDE_TC_TMA_STORE=0 IRSDE_PROFILE_DUMP=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/bench_prof_nostore.log 2> gpurun_out/prof_dump_nostore.txt
fi
