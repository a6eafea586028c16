#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_rpl.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_rpl.log
tail -8 gpurun_out/pytest_rpl.log | cut -c1-300
{
for w in k4 k5 k4i; do for m in 0 2; do TGNH_RPL=$m python scripts/dev_rpl_probe.py $w 2>&1 | tail -1; done; done
} | tee gpurun_out/rpl_probe.log
for m in 0 2; do echo "TGNH_RPL=$m"; TGNH_RPL=$m python scripts/dev_halves_time.py 2>&1 | tail -5; done | tee gpurun_out/rpl_halves.log
TGNH_RPL=1 ncu --set full --clock-control none --import-source on -k regex:tgnh_v2 -s 9 -c 1 -o gpurun_out/prof_rpl \
    python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --quick --no-reference-cuda > gpurun_out/ncu_rpl.log 2>&1
ncu -i gpurun_out/prof_rpl.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rr=list(csv.reader(sys.stdin)); h=rr[0]
for r in rr[2:]:
    d=dict(zip(h,r))
    for k in ['Kernel Name','gpu__time_duration.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']:
        print(k, d.get(k))
" | tee gpurun_out/ncu_rpl_metrics.txt
