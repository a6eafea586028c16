#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python scripts/dev_ke_error.py > gpurun_out/dev_ke_error.log 2>&1; cat gpurun_out/dev_ke_error.log | cut -c1-200
