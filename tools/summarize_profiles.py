#!/usr/bin/env python
"""Turn the raw ncu outputs in gpurun_out/ (written by tools/profile.sh on the B200 box) into the
tracked summaries under profiles/:  rNN_launches.csv (per-kernel share of the step), rNN_kernels.md (key ncu
metrics per kernel incl. DRAM traffic) and traffic.json (dram bytes per launch, read by bench.py)."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "gpurun_out")
OUT = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
S_FRAMES, N_BA = 444, 148          # units per launch of the profiled command (bench.py defaults)
os.makedirs(OUT, exist_ok=True)



def kname(full):
    """'void vrf::k_pyr<0>(FrontCfg, ...)' -> 'k_pyr<0>'"""
    n = full.split("(")[0].split("::")[-1].strip()
    return n[5:] if n.startswith("void ") else n


rows = []
with open(os.path.join(GO, "launches.csv")) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
        rows.append((kname(r["Kernel Name"]), v))
agg = collections.OrderedDict()
for k, ns in rows:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += ns
ours = {k: v for k, v in agg.items() if k.startswith("k_")}
tot = sum(v[1] for v in ours.values()) or 1.0
with open(os.path.join(OUT, f"{tag}_launches.csv"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include vrf_timed/   VRF_NVTX=1 python bench.py --steps 4 --warmup 3 --quick\n")
    f.write("# (per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's roofline.kernels, not absolutes)\n")
    f.write("kernel,launches,total_us,avg_us,share\n")
    for k, (n, ns) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k},{n},{ns / 1e3:.1f},{ns / n / 1e3:.2f},{ns / tot:.4f}\n")
    other = sum(v[1] for k, v in agg.items() if not k.startswith("k_"))
    f.write(f"# launches of other kernels (torch set-up copies etc.): {sum(v[0] for k, v in agg.items() if not k.startswith('k_'))}, {other / 1e3:.1f} us\n")

want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg"]
to_bytes = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
traffic = {}
with open(os.path.join(OUT, f"{tag}_kernels.md"), "w") as md:
    md.write(f"# ncu --set full captures ({tag}), one launch per kernel of one bench step, B200, `--clock-control none`\n\n")
    md.write("Command: `tools/profile.sh` (`python bench.py --steps 4 --warmup 3 --quick`: 444 sequences, 148 BA windows per launch).\n"
             "`traffic` = dram__bytes_read.sum + dram__bytes_write.sum of that launch.\n\n")
    for fn in sorted(os.listdir(GO)):
        if not fn.endswith(".ncu-rep") or not fn.startswith("prof_all"):
            continue
        out = subprocess.run(["ncu", "-i", os.path.join(GO, fn), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(out.splitlines()))
        if len(rr) < 3:
            continue
        hdr, units = rr[0], rr[1]
        seen = set()
        for vals in rr[2:]:
            if len(vals) != len(hdr):
                continue
            name = kname(vals[hdr.index("Kernel Name")])
            if name in seen:
                continue
            seen.add(name)
            md.write(f"## {name}\n\n| metric | value | unit |\n|---|---|---|\n")
            dram = 0.0
            for w in want:
                if w in hdr:
                    i = hdr.index(w)
                    md.write(f"| {w} | {vals[i]} | {units[i]} |\n")
                    if w.startswith("dram__bytes"):
                        dram += float(vals[i].replace(",", "")) * to_bytes.get(units[i], 1.0)
            md.write(f"| traffic (read + write) | {dram / 1e6:.3f} | Mbyte |\n\n")
            ba = name.startswith("k_ba")
            if name not in traffic:
                traffic[name] = {"dram_bytes_per_launch": dram, "grid": int(vals[hdr.index("launch__grid_size")].replace(",", "")),
                                 "note": "ncu --set full, python bench.py --steps 4 --warmup 3 --quick (444 frames / 148 BA windows per launch)",
                                 "units_in_capture": N_BA if ba else S_FRAMES, "unit": "BA windows" if ba else "frames"}
if traffic:
    json.dump(traffic, open(os.path.join(OUT, "traffic.json"), "w"), indent=1, sort_keys=True)
print("wrote", OUT, sorted(traffic))
