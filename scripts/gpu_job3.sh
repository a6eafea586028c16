#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest3.log
tail -30 gpurun_out/pytest3.log
for occ in 0 2 3 1; do
  TGNH_TUNE_OCC=$occ timeout -s KILL 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --quick > gpurun_out/bench_occ$occ.json 2> gpurun_out/bench_occ$occ.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_occ$occ.json")); r=d["roofline"]
    print("occ$occ", round(d["ms_per_step"]*1e3,1), "us/step  A", round(r["avg_launch_ms"]*1e3,1), " B", round(r["half2_avg_launch_ms"]*1e3,1))
except Exception as e: print("occ$occ failed", e)
PY
done
timeout -s KILL 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "full rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_full.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','reference_cuda') if k in d})
print(json.dumps(d['config'],indent=0)[:3000])"
tail -5 gpurun_out/bench_full.err
