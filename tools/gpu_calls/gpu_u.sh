#!/bin/bash
# round 2, call U: back-end handles per publish-phase group (1 vs 3 streams) on the three workloads
cd /root/repo
mkdir -p gpurun_out
for cfgn in c3 c4 c5; do
  for nbs in 1 3; do
    timeout 400 python bench.py --config $cfgn --quick --steps 100 --ba-streams $nbs > gpurun_out/u_bench_${cfgn}_s$nbs.json 2> gpurun_out/u_bench_${cfgn}_s$nbs.err
    python - <<PY
import json
try:
    j=json.load(open("gpurun_out/u_bench_${cfgn}_s$nbs.json"))
    print("$cfgn ba-streams $nbs value", round(j["value"]), "ms/step", round(j["ms_per_step"],3), {k: round(v["ms_per_step"],3) for k,v in j["roofline"]["kernels"].items() if k.startswith("k_ba") or k=="k_lk"})
except Exception as e:
    print("$cfgn $nbs failed", e)
PY
  done
done
tail -3 gpurun_out/u_bench_c3_s3.err | cut -c1-300
