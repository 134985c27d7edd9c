#!/bin/bash
# round 2, call J: parity after RANSAC / BA-pack / LK-reflect changes + e2e pipeline diagnosis
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j_pytest.log
tail -4 gpurun_out/j_pytest.log | cut -c1-300
for mode in "both" "front" "front_copyonly" "front_kernelsonly" "ba"; do
  case $mode in
    both) env="";;
    front) env="VRF_E2E_PART=front";;
    front_copyonly) env="VRF_E2E_PART=front VRF_DEBUG_SKIP=2";;
    front_kernelsonly) env="VRF_E2E_PART=front VRF_DEBUG_SKIP=1";;
    ba) env="VRF_E2E_PART=ba";;
  esac
  env $env timeout 300 python bench.py --quick --with-e2e --steps 20 > gpurun_out/j_e2e_$mode.json 2> gpurun_out/j_e2e_$mode.err
  python - <<PY
import json
j=json.load(open("gpurun_out/j_e2e_$mode.json"))
print("$mode", "value", round(j["value"]), "e2e", j["e2e"]["value"] and round(j["e2e"]["value"]), "gray", j["e2e_gray8"]["value"] and round(j["e2e_gray8"]["value"]), {k: round(v["ms_per_step"],3) for k,v in j["roofline"]["kernels"].items() if k in ("k_ransac","k_lk","k_ba_solve")})
PY
  tail -1 gpurun_out/j_e2e_$mode.err | cut -c1-200
done
