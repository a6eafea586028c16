#!/bin/bash
mkdir -p gpurun_out
./gpurun_variants/chain_bench_new
timeout -s KILL 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mixed.py -m gpu -q -x 2>&1 | tail -4
for lib in gpurun_variants/libtgnh_old.so openmm_drudenose_b200/libtgnh.so; do
for w in c1 c2 c3; do
  TGNH_LIB=$PWD/$lib python bench.py --steps 400 --warmup 10 --workload $w --no-cpu-baseline --no-e2e --quick 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib $w', round(d['ms_per_step']*1e3,2), 'us/step')"
done; done
bash scripts/gpu_job_variants.sh gpurun_variants/libtgnh_old.so openmm_drudenose_b200/libtgnh.so
