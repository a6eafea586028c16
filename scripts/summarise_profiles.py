"""Turns gpurun_out/ artefacts (scripts/collect_profiles.sh) into the tracked summaries under profiles/."""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

R = sys.argv[1] if len(sys.argv) > 1 else "r1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

for name in (f"bench_{R}.json", f"bench_ref_{R}.json", f"bench_c4-wall_{R}.json", f"bench_c4-hot_{R}.json", f"bench_gen1_{R}.json", f"bench_c1_{R}.json", f"bench_c2_{R}.json", f"bench_c3_{R}.json",
             f"launches_{R}.csv", f"gpu_{R}.csv", f"rpl_probe_{R}.log"):
    if os.path.exists(os.path.join(G, name)):
        shutil.copy(os.path.join(G, name), os.path.join(P, name))

# launch list -> per-kernel table
if not os.path.exists(os.path.join(G, f"launches_{R}.csv")):
    rows = []
else:
    rows = [r for r in csv.reader(open(os.path.join(G, f"launches_{R}.csv"))) if len(r) > 5]
agg = collections.OrderedDict()
if rows:
    hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    for r in rows[1:]:
        agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")) / 1000.0)
total = sum(sum(v) for v in agg.values())
lines = [f"# ncu launch list ({R}): `python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline`",
         "# gpu__time_duration.sum per launch, --clock-control none (cold cache, serialised: compare SHARES)", "",
         "| kernel | launches | mean us | min us | max us | share of listed time |", "|---|---|---|---|---|---|"]
for k, v in agg.items():
    lines.append(f"| `{k}` | {len(v)} | {sum(v)/len(v):.2f} | {min(v):.2f} | {max(v):.2f} | {100*sum(v)/total:.1f} % |")
if rows:
    open(os.path.join(P, f"launch_summary_{R}.md"), "w").write("\n".join(lines) + "\n")

# full capture -> raw metrics of interest
rep = os.path.join(G, f"prof_{R}.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, u = rr[0], rr[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
    idx = {x: i for i, x in enumerate(h)}
    out = [f"# ncu --set full --clock-control none ({R}), 10M-particle C4 workload; default cache control (L2 flushed before each pass)", ""]
    summary = {}
    for r in rr[2:]:
        out.append(f"## {r[idx['Kernel Name']]}")
        for w in want[1:]:
            if w in idx:
                out.append(f"- {w}: {r[idx[w]]} {u[idx[w]]}")
        out.append("")
        kn = r[idx["Kernel Name"]]
        kind = "half1" if "<0," in kn else ("half2" if ("<3," in kn or "<1," in kn) else "reduce")
        def num(m):
            return float(r[idx[m]].replace(",", ""))
        ur, uw = u[idx["dram__bytes_read.sum"]], u[idx["dram__bytes_write.sum"]]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        summary[kind] = {"kernel": r[idx["Kernel Name"]], "dram_bytes": num("dram__bytes_read.sum") * scale[ur] + num("dram__bytes_write.sum") * scale[uw],
                         "duration_us": num("gpu__time_duration.sum")}
    open(os.path.join(P, f"ncu_full_{R}.md"), "w").write("\n".join(out))
    json.dump(summary, open(os.path.join(P, f"ncu_full_{R}.json"), "w"), indent=1)
print("profiles/ updated:", sorted(os.listdir(P)))
