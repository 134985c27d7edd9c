#!/bin/bash
# round 2, call L: deterministic BA accumulation (no FP64 atomics) + k_pyr two-items-in-flight; parity, bench, ncu of k_pyr
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/l_pytest.log
tail -8 gpurun_out/l_pytest.log | cut -c1-400
timeout 300 python bench.py --quick --steps 30 > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err
python - <<PY
import json
j=json.load(open("gpurun_out/l_bench.json"))
print("value", round(j["value"]), {k: round(v["ms_per_step"],4) for k,v in j["roofline"]["kernels"].items()}, j["roofline"].get("image_scan_kernels"))
print(j["roofline"]["ba_solve_phase_cycles"])
PY
grep "ba slot" gpurun_out/l_bench.err | head -8 | cut -c1-330
export VRF_NVTX=1
timeout 400 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "vrf_profile_step/" -k regex:k_pyr -c 2 -f -o gpurun_out/l_prof_k_pyr python bench.py --steps 3 --warmup 3 --quick > gpurun_out/l_prof_k_pyr.log 2>&1
ls -la gpurun_out | grep l_prof
