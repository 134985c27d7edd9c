#!/bin/bash
# round 2, call X: four landmarks per pass in the W-row traversals of k_ba_solve: parity + quick bench c3 / c4
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "ba or golden or edge" > gpurun_out/x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/x_pytest.log
tail -3 gpurun_out/x_pytest.log | cut -c1-300
for c in c3 c4; do
  timeout 300 python bench.py --config $c --quick --steps 60 > gpurun_out/x_bench_$c.json 2> gpurun_out/x_bench_$c.err
  python - <<PY
import json
j=json.load(open("gpurun_out/x_bench_$c.json"))
print("$c value", round(j["value"]), {k: round(v["ms_per_step"],3) for k,v in j["roofline"]["kernels"].items() if k.startswith("k_ba")}, j["roofline"]["ba_solve_phase_cycles"])
PY
done
