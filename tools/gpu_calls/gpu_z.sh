#!/bin/bash
# round 2, call Z: final tree -- the default bench line and the reference arm
cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/z_bench_c3.json 2> gpurun_out/z_bench_c3.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 --ref-frames 24 > gpurun_out/z_bench_ref_c3.json 2> gpurun_out/z_bench_ref_c3.err
python - <<PY
import json
j=json.load(open("gpurun_out/z_bench_c3.json")); r=json.load(open("gpurun_out/z_bench_ref_c3.json"))
print("c3 value", round(j["value"]), "ms", round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"]), "gray", round(j["e2e_gray8"]["value"]), "cpu", round(j["cpu_baseline"]["value"]), "ref arm", round(r["value"]))
PY
