#!/bin/bash
# Run on the GPU box (gpurun): (1) kernel launch list of the timed region of the bench command, (2) one `ncu --set full`
# capture of every kernel of one bench step (444 sequences / 148 BA windows per launch, as in the bench line), (3) source-level
# captures of the two kernels that carry the path.  bench.py marks the regions with NVTX ranges when VRF_NVTX is set.
# Outputs land in gpurun_out/ ; tools/summarize_profiles.py turns them into the tracked summaries under profiles/.
mkdir -p gpurun_out
export VRF_NVTX=1
CMD="python bench.py --steps 4 --warmup 3 --quick"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "vrf_timed/" --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --nvtx --nvtx-include "vrf_profile_step/" -k regex:^k_ -f -o gpurun_out/prof_all $CMD > gpurun_out/prof_all.log 2>&1
CMD2="python bench.py --steps 3 --warmup 3 --seqs 96 --quick"
for k in ${KERNELS:-k_lk k_ba_solve}; do
  timeout 400 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "vrf_profile_step/" -k regex:$k -c 1 -f -o gpurun_out/prof_$k $CMD2 > gpurun_out/prof_$k.log 2>&1
done
ls -la gpurun_out | grep -E "prof_|launches"
# afterwards, here:  python tools/summarize_profiles.py r02 ;  for k in k_lk k_ba_solve; do ncu -i gpurun_out/prof_$k.ncu-rep --page source
#   --csv --print-source cuda,sass > /tmp/$k.csv; python tools/ncu_source_summary.py /tmp/$k.csv 45 > profiles/r02_hotlines_$k.txt; done
