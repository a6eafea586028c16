#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests/test_refcuda.py -m gpu -x -q -s > gpurun_out/refcuda.log 2>&1; echo "refcuda rc=$?" >> gpurun_out/refcuda.log
tail -25 gpurun_out/refcuda.log
timeout -s KILL 600 python -m pytest tests/test_plugin.py -m gpu -x -q > gpurun_out/plugin.log 2>&1; tail -3 gpurun_out/plugin.log
bash scripts/gpu_ncu.sh r2a
