#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest4.log
tail -40 gpurun_out/pytest4.log | cut -c1-300
for t in 0 1 2 3 4 8; do
  TGNH_TUNE_B=$t timeout -s KILL 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --quick > gpurun_out/bench_t$t.json 2> gpurun_out/bench_t$t.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_t$t.json")); r=d["roofline"]
    print("tune$t", round(d["ms_per_step"]*1e3,1), "us/step  A", round(r["avg_launch_ms"]*1e3,1), " B", round(r["half2_avg_launch_ms"]*1e3,1))
except Exception as e: print("tune$t failed", e)
PY
done
