"""Warp-stall sampling of the two hot kernels from the round's `ncu --set full` capture -> profiles/ncu_stalls_<round>.md"""
import csv, io, os, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = os.path.join(ROOT, "gpurun_out", f"prof_{R}.ncu-rep")
out = [f"# Warp-stall sampling per SASS instruction ({R}; `ncu --set full --import-source on`, C4 workload)", ""]
for skip, title in ((0, "first-half kernel"), (1, "second-half kernel")):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(skip), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    kname = rr[0][1] if rr and len(rr[0]) > 1 else "?"
    h = rr[1]
    rows = [r for r in rr[2:] if len(r) > 45 and r[2].isdigit()]
    seen, uniq = set(), []
    for r in rows:                      # the export lists every row twice
        if r[0] not in seen:
            seen.add(r[0]); uniq.append(r)
    tot = sum(int(r[2]) for r in uniq)
    out += [f"## {title}: `{kname}`", "", f"{tot} samples over {len(uniq)} instructions", "", "| stall reason | samples | share |", "|---|---|---|"]
    cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    tots = sorted(((sum(int(r[i] or 0) for r in uniq), h[i]) for i in cols), reverse=True)
    for v, name in tots:
        if v:
            out.append(f"| {name} | {v} | {100.0 * v / tot:.1f} % |")
    out += ["", "| samples | instruction | dominant reasons |", "|---|---|---|"]
    for r in sorted(uniq, key=lambda r: -int(r[2]))[:12]:
        why = ", ".join(f"{h[i][6:]} {int(r[i])}" for i in cols if int(r[i] or 0) > 0.15 * int(r[2]))
        out.append(f"| {r[2]} | `{r[1].strip()[:70]}` | {why} |")
    out.append("")
open(os.path.join(ROOT, "profiles", f"ncu_stalls_{R}.md"), "w").write("\n".join(out))
print("\n".join(out[:40]))
