#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_r2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2.log
tail -30 gpurun_out/pytest_r2.log | cut -c1-400
