#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_refcuda.py tests/test_gpu_mixed.py -m gpu -q > gpurun_out/pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest5.log
grep -n "^E  \|passed\|failed\|^FAILED" gpurun_out/pytest5.log | cut -c1-250 | tail -30
timeout -s KILL 600 python scripts/dev_refcuda_1000.py > gpurun_out/dev_refcuda_1000.log 2>&1; tail -5 gpurun_out/dev_refcuda_1000.log
for lib in libtgnh.so libtgnh_w31.so; do
  TGNH_LIB=$PWD/openmm_drudenose_b200/$lib timeout -s KILL 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --quick > gpurun_out/bench_$lib.json 2> gpurun_out/bench_$lib.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$lib.json")); r=d["roofline"]
    print("$lib", round(d["ms_per_step"]*1e3,1), "us/step  A", round(r["avg_launch_ms"]*1e3,1), " B", round(r["half2_avg_launch_ms"]*1e3,1))
except Exception as e: print("$lib failed", e)
PY
done
TGNH_LIB=$PWD/openmm_drudenose_b200/libtgnh_w31.so timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
