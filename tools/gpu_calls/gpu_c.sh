#!/bin/bash
# round 2, call C: full GPU suite, BA evaluate breakdown, ncu source capture of the new k_lk and of k_ba_solve, default bench line
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
tail -3 gpurun_out/c_pytest.log
VRF_BA_DEBUG=16 timeout 120 python bench.py --quick --seqs 6 --steps 3 --warmup 3 > gpurun_out/c_badebug.json 2> gpurun_out/c_badebug.err
grep -c evaluate gpurun_out/c_badebug.err
CMD="python bench.py --steps 3 --warmup 3 --seqs 96 --quick"
for k in k_lk k_ba_solve; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 1 -f -o gpurun_out/c_prof_$k $CMD > gpurun_out/c_prof_$k.log 2>&1
done
timeout 600 python bench.py > gpurun_out/c_bench_c3.json 2> gpurun_out/c_bench_c3.err
tail -2 gpurun_out/c_bench_c3.err
ls -la gpurun_out | tail -8
