#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
for w in c1 c2 c3; do
  python bench.py --steps 200 --warmup 10 --workload $w --no-cpu-baseline --no-e2e --quick > gpurun_out/small_$w.json 2> gpurun_out/small_$w.err
  python -c "
import json; d=json.load(open('gpurun_out/small_$w.json')); print('$w', round(d['ms_per_step']*1e3,2), 'us/step')"
done
bash scripts/gpu_job_variants.sh openmm_drudenose_b200/libtgnh.so
