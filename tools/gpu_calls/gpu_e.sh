#!/bin/bash
# round 2, call E: DMMA trailing update + unscaled W: parity suite, bench, ncu (k_lk warm launch, launch list, k_ba_solve)
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
tail -3 gpurun_out/e_pytest.log
timeout 300 python bench.py --quick --steps 30 > gpurun_out/e_bench_c3.json 2> gpurun_out/e_bench_c3.err
tail -9 gpurun_out/e_bench_c3.err | cut -c1-330
CMD="python bench.py --steps 3 --warmup 3 --seqs 96 --quick"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_lk -s 5 -c 1 -f -o gpurun_out/e_prof_k_lk $CMD > gpurun_out/e_prof_k_lk.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_ba_solve -s 100 -c 1 -f -o gpurun_out/e_prof_k_ba_solve $CMD > gpurun_out/e_prof_k_ba_solve.log 2>&1
tail -2 gpurun_out/e_prof_k_ba_solve.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 3000 --csv --log-file gpurun_out/e_launches.csv $CMD > gpurun_out/e_launches.log 2>&1
grep -c k_lk gpurun_out/e_launches.csv
ls -la gpurun_out | grep " e_"
