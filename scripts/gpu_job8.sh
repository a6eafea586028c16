#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python scripts/dev_bias2.py > gpurun_out/dev_bias.log 2>&1; cat gpurun_out/dev_bias.log | cut -c1-250
