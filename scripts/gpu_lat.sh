#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -k "latent or nafnet" -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/t_lat.log
tail -n 30 gpurun_out/t_lat.log
