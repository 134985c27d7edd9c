#!/bin/bash
# round 2, call N: projected Armijo line search in k_ba_solve (parity vs oracle) + bench
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n_pytest.log
tail -12 gpurun_out/n_pytest.log | cut -c1-600
timeout 300 python bench.py --quick --steps 30 > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err
python - <<PY
import json
j=json.load(open("gpurun_out/n_bench.json"))
print("value", round(j["value"]), {k: round(v["ms_per_step"],4) for k,v in j["roofline"]["kernels"].items()})
print(j["roofline"]["ba_solve_phase_cycles"])
PY
grep "ba slot" gpurun_out/n_bench.err | head -8 | cut -c1-330
