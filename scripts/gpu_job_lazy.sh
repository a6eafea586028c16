#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lazy or single_step or full_size or thousand or half_calls" 2>&1 | tail -5
for lz in 1 0 1; do
TGNH_LAZY_KICK=$lz timeout -s KILL 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --quick --no-reference-cuda > gpurun_out/bench_lazy$lz.json 2> gpurun_out/bench_lazy$lz.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_lazy$lz.json")); r=d["roofline"]
    print("lazy=$lz", round(d["ms_per_step"]*1e3,1), "us/step  A", round(r["avg_launch_ms"]*1e3,1), " B", round(r["half2_avg_launch_ms"]*1e3,1))
except Exception as e: print("lazy=$lz failed", e)
PY
done
