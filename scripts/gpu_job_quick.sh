#!/bin/bash
# quick parity subset + headline bench of the listed builds
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "single_step or full_size or thousand or big_res or ragged" 2>&1 | tail -3
bash scripts/gpu_job_variants.sh "$@"
