#!/bin/bash
# residue-per-lane reduction: parity suite, then headline bench + call-pattern timings with TGNH_RPL = 0 / 1 / 2
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_rpl.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_rpl.log
tail -8 gpurun_out/pytest_rpl.log | cut -c1-400
for lib in openmm_drudenose_b200/libtgnh.so gpurun_variants/libtgnh_ns6.so; do
  [ -f $lib ] || continue
  for m in 0 1 2; do
    n=$(basename $lib .so)_rpl$m
    TGNH_RPL=$m TGNH_LIB=$PWD/$lib timeout -s KILL 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --quick --no-reference-cuda > gpurun_out/bench_$n.json 2> gpurun_out/bench_$n.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$n.json")); r=d["roofline"]
    print("$n", round(d["ms_per_step"]*1e3,1), "us/step  A", round(r["avg_launch_ms"]*1e3,1), " B", round(r["half2_avg_launch_ms"]*1e3,1))
except Exception as e: print("$n failed", e)
PY
  done
done
for m in 0 2; do echo "TGNH_RPL=$m"; TGNH_RPL=$m python scripts/dev_halves_time.py 2>&1 | tail -5; done
