#!/bin/bash
# round 2, last call: host-buffer arm of c3 with 1 vs 2 back-end streams (same box, back to back)
cd /root/repo
mkdir -p gpurun_out
for nbs in 1 2; do
  timeout 200 python bench.py --quick --with-e2e --steps 20 --ba-streams $nbs > gpurun_out/e2_bench_s$nbs.json 2> gpurun_out/e2_bench_s$nbs.err
  python - <<PY
import json
j=json.load(open("gpurun_out/e2_bench_s$nbs.json"))
print("ba-streams $nbs value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "gray", round(j["e2e_gray8"]["value"]))
PY
done
