#!/bin/bash
echo "== old"; TGNH_LIB=$PWD/gpurun_variants/libtgnh_old.so python scripts/dev_chain_time.py 2>&1 | tail -4
echo "== new"; python scripts/dev_chain_time.py 2>&1 | tail -4
for lib in gpurun_variants/libtgnh_old.so openmm_drudenose_b200/libtgnh.so; do
TGNH_LIB=$PWD/$lib python bench.py --steps 400 --warmup 10 --workload c1 --no-cpu-baseline --no-e2e --quick 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib c1', round(d['ms_per_step']*1e3,2), 'us/step')"
done
