#!/bin/bash
# round 2, call S: compute-sanitizer (memcheck + racecheck) over the rebuilt kernels, then smoke()
cd /root/repo
mkdir -p gpurun_out
for t in memcheck racecheck; do
  for c in ba ba_ex front; do
    timeout 600 compute-sanitizer --tool $t --print-limit 20 python tools/scratch/sanitize_case.py $c > gpurun_out/s_${t}_$c.log 2>&1
    echo "$t $c: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/s_${t}_$c.log | tail -1)"
  done
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s_smoke.log 2>&1; tail -1 gpurun_out/s_smoke.log
