#!/bin/bash
# round 2, call G: pair-major factor lists + in-register diag8: parity + bench + Cholesky phase timing
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_edge_gpu.py tests/test_golden.py tests/test_fm_gpu.py -m gpu -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g_pytest.log
tail -6 gpurun_out/g_pytest.log | cut -c1-300
VRF_BA_DEBUG=32 timeout 120 python bench.py --quick --seqs 3 --steps 3 --warmup 3 > gpurun_out/g_badebug.json 2> gpurun_out/g_badebug.err
grep "chol warp" gpurun_out/g_badebug.err | tail -8
timeout 300 python bench.py --quick --steps 30 > gpurun_out/g_bench_c3.json 2> gpurun_out/g_bench_c3.err
grep "ba slot" gpurun_out/g_bench_c3.err | head -3 | cut -c1-200
