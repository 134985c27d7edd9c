#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/w_bench_c3.json 2> gpurun_out/w_bench_c3.err
python - <<PY
import json
j=json.load(open("gpurun_out/w_bench_c3.json"))
print("c3 value", round(j["value"]), "ms", round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"]), "gray", round(j["e2e_gray8"]["value"]), "cpu", round(j["cpu_baseline"]["value"]), j["config"]["ba_streams"])
PY
