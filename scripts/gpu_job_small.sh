#!/bin/bash
# per-launch device time of the small-system kernels (ncu launch list; cold-ish caches)
mkdir -p gpurun_out
for w in c1 c2 c3; do
  python bench.py --steps 200 --warmup 10 --workload $w --no-cpu-baseline --no-e2e --quick > gpurun_out/small_$w.json 2> gpurun_out/small_$w.err
  ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/small_launches_$w.csv \
     python bench.py --steps 20 --warmup 10 --workload $w --no-cpu-baseline --no-e2e --quick > gpurun_out/small_ncu_$w.log 2>&1
  python - <<PY
import csv, json, collections
d=json.load(open("gpurun_out/small_$w.json")); print("$w", round(d["ms_per_step"]*1e3,2), "us/step")
rows=[r for r in csv.reader(open("gpurun_out/small_launches_$w.csv")) if len(r)>5]
h=rows[0]; ki,vi=h.index("Kernel Name"),h.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[1:]: agg.setdefault(r[ki][:70],[]).append(float(r[vi].replace(",",""))/1000)
for k,v in agg.items(): print("   ",k,len(v),"x mean %.2f us min %.2f"%(sum(v)/len(v),min(v)))
PY
done
