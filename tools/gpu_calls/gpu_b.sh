#!/bin/bash
# round 2, call B: new k_lk (TMA-staged tiles, IDP.2A interpolation): bit-exact parity + timing
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_frontend_gpu.py tests/test_golden.py -m gpu -x -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -5 gpurun_out/b_pytest.log
timeout 300 python bench.py --config c3 --quick --steps 30 > gpurun_out/b_bench_c3.json 2> gpurun_out/b_bench_c3.err
tail -2 gpurun_out/b_bench_c3.err
