#!/bin/bash
# round 2, call F: fixed DMMA trailing update: BA parity tests + bench
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_edge_gpu.py tests/test_golden.py tests/test_fm_gpu.py -m gpu -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
tail -12 gpurun_out/f_pytest.log | cut -c1-300
timeout 300 python bench.py --quick --steps 30 > gpurun_out/f_bench_c3.json 2> gpurun_out/f_bench_c3.err
grep "ba slot" gpurun_out/f_bench_c3.err | head -3 | cut -c1-200
