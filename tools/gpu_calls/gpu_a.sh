#!/bin/bash
# round 2, call A: parity at the new config sizes + first bench lines of c4 / c5 (no CPU baseline)
cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 300 python bench.py --config c3 --no-cpu-baseline --steps 40 > gpurun_out/a_bench_c3.json 2> gpurun_out/a_bench_c3.err
timeout 300 python bench.py --config c4 --no-cpu-baseline --steps 40 > gpurun_out/a_bench_c4.json 2> gpurun_out/a_bench_c4.err
timeout 300 python bench.py --config c5 --no-cpu-baseline --steps 40 > gpurun_out/a_bench_c5.json 2> gpurun_out/a_bench_c5.err
tail -3 gpurun_out/a_pytest.log
