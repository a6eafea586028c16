#!/bin/bash
# usage: gpu_job_variants.sh <lib> [<lib> ...]  — headline bench of experimental builds of libtgnh.so (TGNH_LIB)
mkdir -p gpurun_out
for lib in "$@"; do
  n=$(basename $lib .so)
  TGNH_LIB=$PWD/$lib timeout -s KILL 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --quick --no-reference-cuda > gpurun_out/bench_$n.json 2> gpurun_out/bench_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$n.json")); r=d["roofline"]
    print("$n", round(d["ms_per_step"]*1e3,1), "us/step  A", round(r["avg_launch_ms"]*1e3,1), " B", round(r["half2_avg_launch_ms"]*1e3,1))
except Exception as e: print("$n failed", e)
PY
done
