#!/bin/bash
# round 2, call I: the bench lines of the three workloads (full: CPU baseline + e2e) and the ncu captures for profiles/
cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/i_bench_c3.json 2> gpurun_out/i_bench_c3.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 --ref-frames 24 > gpurun_out/i_bench_ref_c3.json 2> gpurun_out/i_bench_ref_c3.err
timeout 900 python bench.py --config c4 > gpurun_out/i_bench_c4.json 2> gpurun_out/i_bench_c4.err
timeout 900 python bench.py --config c5 > gpurun_out/i_bench_c5.json 2> gpurun_out/i_bench_c5.err
bash tools/profile.sh > gpurun_out/i_profile.log 2>&1
tail -5 gpurun_out/i_profile.log
for c in c3 c4 c5; do head -c 300 gpurun_out/i_bench_$c.json; echo; done
