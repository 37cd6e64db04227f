#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -k "latent" -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/t_lat.log
timeout 300 python scripts/time_latent.py 1 1024 > gpurun_out/lat_b1.log 2>&1
timeout 300 python scripts/time_latent.py 4 1024 > gpurun_out/lat_b4.log 2>&1
tail -n 3 gpurun_out/t_lat.log; tail -n 5 gpurun_out/lat_b1.log | cut -c1-1500; tail -n 5 gpurun_out/lat_b4.log | cut -c1-1500
