#!/bin/bash
mkdir -p gpurun_out
{
python scripts/dev_small_probe.py c2
TGNH_V2=0 python scripts/dev_small_probe.py c2
python scripts/dev_small_probe.py c3
TGNH_FUSE_CHAIN=0 python scripts/dev_small_probe.py c3
TGNH_V2=0 python scripts/dev_small_probe.py c1
} > gpurun_out/small_probe.log 2>&1
cat gpurun_out/small_probe.log
