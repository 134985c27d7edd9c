#!/bin/bash
# Run on the GPU box (gpurun): ncu --set full + source counters of the two back-end kernels and k_lk.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --seqs 96 --quick"
for k in ${KERNELS:-k_ba_marg k_ba_solve k_lk}; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s ${SKIP:-12} -c 1 -f -o gpurun_out/prof_$k $CMD > gpurun_out/prof_$k.log 2>&1
done
ls -la gpurun_out
# afterwards, here:  for k in ...; do ncu -i gpurun_out/prof_$k.ncu-rep --page source --csv --print-source cuda,sass > /tmp/$k.csv;
#                    python tools/ncu_source_summary.py /tmp/$k.csv 45 > profiles/rNN_hotlines_$k.txt; done
