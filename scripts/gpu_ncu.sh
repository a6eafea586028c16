#!/bin/bash
# one gpurun call: plain run of the profiled command, then one --set full capture of a first-half and a second-half launch
R=${1:-r2a}
mkdir -p gpurun_out
python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --quick > gpurun_out/plain_$R.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tgnh_v2 -s 8 -c 2 -o gpurun_out/prof_$R \
    python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --quick > gpurun_out/ncu_full_$R.log 2>&1
tail -3 gpurun_out/ncu_full_$R.log
