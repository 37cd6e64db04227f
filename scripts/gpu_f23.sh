#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -k "imaging or philox or restorer or nafnet or latent or sharded" -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/t_f23.log
tail -n 30 gpurun_out/t_f23.log | cut -c1-300
timeout 300 python scripts/time_latent.py 1 1024 2>&1 | tail -1 | cut -c1-900
