#!/bin/bash
./gpurun_variants/chain_bench_new
for l in openmm_drudenose_b200/libtgnh.so gpurun_variants/libtgnh_old.so; do echo $l; S_LIST=1,20 TGNH_LIB=$PWD/$l python scripts/dev_c4_chain_exposed.py 2>&1 | grep "^S="; done
for lib in gpurun_variants/libtgnh_old.so openmm_drudenose_b200/libtgnh.so; do
for w in c1 c2 c3; do
  TGNH_LIB=$PWD/$lib python bench.py --steps 400 --warmup 10 --workload $w --no-cpu-baseline --no-e2e --quick 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib $w', round(d['ms_per_step']*1e3,2), 'us/step')"
done; done
