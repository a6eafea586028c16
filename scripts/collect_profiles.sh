#!/bin/bash
# Runs on the GPU box (under gpurun): bench lines, ncu launch list of the same command, one --set full capture of the
# dominant kernel.  Outputs land in gpurun_out/ and are summarised into profiles/ by scripts/summarise_profiles.py.
set -u
R=${1:-r1}
mkdir -p gpurun_out
python bench.py --steps 50 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err || exit 1
python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_ref_$R.json 2>> gpurun_out/bench_$R.err
for w in c4-wall c4-hot; do python bench.py --steps 50 --warmup 3 --workload $w --no-cpu-baseline --no-e2e --quick > gpurun_out/bench_${w}_$R.json 2>> gpurun_out/bench_$R.err; done
TGNH_V2=0 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e --quick > gpurun_out/bench_gen1_$R.json 2>> gpurun_out/bench_$R.err
for w in c1 c2 c3; do python bench.py --steps 200 --warmup 10 --workload $w --no-cpu-baseline --no-e2e --quick > gpurun_out/bench_${w}_$R.json 2>> gpurun_out/bench_$R.err; done
# launch list (per-launch device time, cold cache, serialised) of the same command as the bench
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --quick > gpurun_out/plain_$R.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --quick > gpurun_out/ncu_list_$R.log 2>&1
# the dominant kernel (first half) and the second-half kernel, full set
ncu --set full --clock-control none --import-source on -k regex:tgnh_v2 -s 8 -c 2 -o gpurun_out/prof_$R \
    python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --quick > gpurun_out/ncu_full_$R.log 2>&1
# residue-per-lane reduction on / off, 10 M particles: 4-site box (C4's molecules) and 5-site SWM4 box (massless M site)
{ for w in k4 k5 k4i; do for m in 0 2; do TGNH_RPL=$m python scripts/dev_rpl_probe.py $w 2>&1 | tail -1; done; done; } > gpurun_out/rpl_probe_$R.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > gpurun_out/gpu_$R.csv
echo collected
