#!/bin/bash
mkdir -p gpurun_out
for t in 0 0.3 0.45 0.6 0.7 0.8 1.0; do
  TGNH_TUNE_L2=$t timeout -s KILL 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --quick > gpurun_out/bench_l2_$t.json 2> gpurun_out/bench_l2_$t.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_l2_$t.json")); r=d["roofline"]
    print("keep $t", round(d["ms_per_step"]*1e3,1), "us/step  A", round(r["avg_launch_ms"]*1e3,1), " B", round(r["half2_avg_launch_ms"]*1e3,1))
except Exception as e: print("keep $t failed", e)
PY
done
