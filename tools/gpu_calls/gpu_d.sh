#!/bin/bash
# round 2, call D: new tests (LMedS, bounds), BA per-task timing, ncu launch list + k_lk capture, c3 quick bench with the new BA inputs
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
tail -3 gpurun_out/d_pytest.log
VRF_BA_DEBUG=16 timeout 120 python bench.py --quick --seqs 3 --steps 3 --warmup 3 > gpurun_out/d_badebug.json 2> gpurun_out/d_badebug.err
grep -c "task lin" gpurun_out/d_badebug.err
timeout 300 python bench.py --quick --steps 30 > gpurun_out/d_bench_c3.json 2> gpurun_out/d_bench_c3.err
tail -9 gpurun_out/d_bench_c3.err | cut -c1-400
CMD="python bench.py --steps 3 --warmup 3 --seqs 96 --quick"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/d_launches.csv $CMD > gpurun_out/d_launches.log 2>&1
grep -c k_lk gpurun_out/d_launches.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_lk -c 1 -f -o gpurun_out/d_prof_k_lk $CMD > gpurun_out/d_prof_k_lk.log 2>&1
tail -3 gpurun_out/d_prof_k_lk.log
ls -la gpurun_out | grep " d_"
