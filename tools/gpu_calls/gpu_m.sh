#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m_pytest.log
tail -8 gpurun_out/m_pytest.log | cut -c1-400
timeout 300 python bench.py --quick --steps 30 > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err
python - <<PY
import json
j=json.load(open("gpurun_out/m_bench.json"))
print("value", round(j["value"]), {k: round(v["ms_per_step"],4) for k,v in j["roofline"]["kernels"].items()}, j["roofline"].get("image_scan_kernels"))
print(j["roofline"]["ba_solve_phase_cycles"])
PY
grep "ba slot" gpurun_out/m_bench.err | head -8 | cut -c1-330
