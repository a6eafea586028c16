#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_r2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2.log
tail -5 gpurun_out/pytest_r2.log | cut -c1-400
for w in k5 k4; do python scripts/dev_rpl_probe.py $w 2>&1 | tail -1; done | tee gpurun_out/rpl_probe2.log
