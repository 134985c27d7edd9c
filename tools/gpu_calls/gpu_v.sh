#!/bin/bash
# round 2, call V: full bench lines (CPU baseline + e2e) of the three workloads with one back-end stream per publish-phase group
cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/v_bench_c3.json 2> gpurun_out/v_bench_c3.err
timeout 900 python bench.py --config c4 > gpurun_out/v_bench_c4.json 2> gpurun_out/v_bench_c4.err
timeout 900 python bench.py --config c5 > gpurun_out/v_bench_c5.json 2> gpurun_out/v_bench_c5.err
python - <<PY
import json
for c in ("c3","c4","c5"):
    try:
        j=json.load(open("gpurun_out/v_bench_%s.json" % c))
        print(c, "value", round(j["value"]), "ms", round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"]), "gray", round(j["e2e_gray8"]["value"]), "cpu", round(j["cpu_baseline"]["value"]), "launches", j["gpu_launches"])
    except Exception as e:
        print(c, "failed", e)
PY
tail -2 gpurun_out/v_bench_c4.err | cut -c1-300
