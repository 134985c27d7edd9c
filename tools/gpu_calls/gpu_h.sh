#!/bin/bash
# round 2, call H: blocked back substitution, coalesced Cauchy, 2-tile DMMA trailing update: full parity suite + benches
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h_pytest.log
tail -6 gpurun_out/h_pytest.log | cut -c1-300
VRF_BA_DEBUG=32 timeout 120 python bench.py --quick --seqs 3 --steps 3 --warmup 3 > gpurun_out/h_badebug.json 2> gpurun_out/h_badebug.err
grep "chol warp" gpurun_out/h_badebug.err | tail -4
timeout 300 python bench.py --quick --steps 30 > gpurun_out/h_bench_c3.json 2> gpurun_out/h_bench_c3.err
grep "ba slot" gpurun_out/h_bench_c3.err | head -3 | cut -c1-200
timeout 300 python bench.py --config c4 --quick --with-e2e --steps 30 > gpurun_out/h_bench_c4.json 2> gpurun_out/h_bench_c4.err
timeout 300 python bench.py --config c5 --quick --with-e2e --steps 30 > gpurun_out/h_bench_c5.json 2> gpurun_out/h_bench_c5.err
