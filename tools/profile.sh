#!/bin/bash
# Run on the GPU box (gpurun): kernel launch list + one full ncu capture of the heaviest kernels.
# Outputs land in gpurun_out/ ; summaries are copied to profiles/ by hand.
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 4 --warmup 3 --seqs 96 --quick"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/launches.log 2>&1
for k in k_lk k_ba_solve k_ingest k_pyrdown k_fast k_ba_marg; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_$k $CMD > gpurun_out/prof_$k.log 2>&1
done
ls -la gpurun_out
