#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -k "nafnet or latent or sharded or full_size or bf16" -p no:cacheprovider 2>&1 | tail -4
timeout 300 python scripts/time_latent.py 1 1024 2>&1 | tail -1 | cut -c1-1200
timeout 300 python scripts/time_latent.py 8 1024 2>&1 | tail -1 | cut -c1-1200
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-400
