#!/bin/bash
for i in 1 2 3; do
for w in c3 c1; do
  python bench.py --steps 400 --warmup 10 --workload $w --no-cpu-baseline --no-e2e --quick 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$w', round(d['ms_per_step']*1e3,2), 'us/step')"
done; done
timeout -s KILL 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
