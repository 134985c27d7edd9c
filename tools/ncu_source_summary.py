#!/usr/bin/env python
"""Per-source-line summary of an `ncu --set full --import-source on` capture.

    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > prof.csv
    python tools/ncu_source_summary.py prof.csv [top_n]

Prints the hottest CUDA source lines by warp-stall samples with their dominant stall reasons,
executed instructions and shared-memory wavefront excess (bank conflicts)."""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(path, newline="")))
    hdr = None
    cur_file = "?"
    agg = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Name":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr) or r[0] == "":
            continue
        d = dict(zip(hdr, r))          # note: two "Source" columns; the dict keeps the last (SASS, '-' on source rows)
        try:
            samples = int(d["# Samples"])
            inst = int(d["Instructions Executed"])
        except ValueError:
            continue
        if samples == 0 and inst == 0:
            continue
        stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
        key = (cur_file, int(r[0]))
        a = agg.setdefault(key, {"src": r[1].strip(), "samples": 0, "inst": 0, "stalls": defaultdict(int), "wf": 0, "wf_ideal": 0})
        a["samples"] += samples; a["inst"] += inst
        a["wf"] += int(d.get("L1 Wavefronts Shared", "0") or 0); a["wf_ideal"] += int(d.get("L1 Wavefronts Shared Ideal", "0") or 0)
        for k, v in stalls.items():
            a["stalls"][k] += v
    tot = sum(a["samples"] for a in agg.values()) or 1
    tot_inst = sum(a["inst"] for a in agg.values()) or 1
    print(f"total samples {tot}, warp instructions {tot_inst}")
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = ", ".join(f"{k} {100 * v // max(a['samples'], 1)}%" for k, v in sorted(a["stalls"].items(), key=lambda kv: -kv[1])[:3])
        wf = f" smem-wf {a['wf']}/{a['wf_ideal']}" if a["wf"] else ""
        print(f"{100 * a['samples'] / tot:5.1f}%  inst {100 * a['inst'] / tot_inst:4.1f}%  {f}:{ln:<4d} [{st}]{wf}  | {a['src'][:110]}")


if __name__ == "__main__":
    main()
