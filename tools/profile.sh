#!/bin/bash
# Run on the GPU box (gpurun): (1) kernel launch list of the bench command, (2) one `ncu --set full` capture of every
# kernel of one bench step (same 444 sequences / 148 BA windows per launch as the bench line).
# Outputs land in gpurun_out/ ; tools/summarize_profiles.py turns them into the tracked summaries under profiles/.
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 4 --warmup 3 --quick"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/launches.log 2>&1
# warm-up = 3 steps x 11 launches (+ set-up launches of torch are not counted: -k filters on our kernels)
timeout 900 ncu --set full --clock-control none -k regex:^k_ -s 44 -c 13 -f -o gpurun_out/prof_all $CMD > gpurun_out/prof_all.log 2>&1
ls -la gpurun_out
