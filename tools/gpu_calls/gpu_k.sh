#!/bin/bash
# round 2, call K: fused ingest + pyrDown kernel (k_pyr): parity + strip-height sweep
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/k_pytest.log
tail -5 gpurun_out/k_pytest.log | cut -c1-400
for r1 in 16 8 32; do
  VRF_PYR_R1=$r1 timeout 300 python bench.py --quick --steps 30 > gpurun_out/k_bench_r$r1.json 2> gpurun_out/k_bench_r$r1.err
  python - <<PY
import json
j=json.load(open("gpurun_out/k_bench_r$r1.json"))
print("R1=$r1 value", round(j["value"]), {k: round(v["ms_per_step"],4) for k,v in j["roofline"]["kernels"].items()}, j["roofline"].get("image_scan_kernels"))
PY
done
