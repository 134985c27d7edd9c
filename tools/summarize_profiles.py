#!/usr/bin/env python
"""Turn the raw ncu outputs in gpurun_out/ (written by tools/profile.sh on the B200 box) into the
tracked summaries under profiles/:  rNN_launches.csv (per-kernel share of the step) and
rNN_kernels.md (key ncu metrics per kernel incl. DRAM traffic)."""
import csv
import collections
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "gpurun_out")
OUT = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(OUT, exist_ok=True)

rows = []
with open(os.path.join(GO, "launches.csv")) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"].split("(")[0], float(r["Metric Value"])))
agg = collections.OrderedDict()
for k, ns in rows:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += ns
tot = sum(v[1] for v in agg.values())
with open(os.path.join(OUT, f"{tag}_launches.csv"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares, not absolutes)\n")
    f.write("kernel,launches,total_us,avg_us,share\n")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k},{n},{ns / 1e3:.1f},{ns / n / 1e3:.2f},{ns / tot:.4f}\n")

want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
with open(os.path.join(OUT, f"{tag}_kernels.md"), "w") as md:
    md.write(f"# ncu --set full captures ({tag}), one launch per kernel, B200, `--clock-control none`\n\n")
    md.write("Command: `tools/profile.sh` (bench.py --quick --seqs 96).  `traffic` = dram read + write of that launch.\n\n")
    for fn in sorted(os.listdir(GO)):
        if not fn.endswith(".ncu-rep"):
            continue
        out = subprocess.run(["ncu", "-i", os.path.join(GO, fn), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(out.splitlines()))
        if len(rr) < 3:
            continue
        hdr, units, vals = rr[0], rr[1], rr[2]
        md.write(f"## {fn[5:-8]}\n\n| metric | value | unit |\n|---|---|---|\n")
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                md.write(f"| {w} | {vals[i]} | {units[i]} |\n")
        md.write("\n")
print("wrote", OUT)
