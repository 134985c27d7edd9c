#!/bin/bash
# round 2, call Y: final tree -- whole GPU suite + smoke()
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/y_pytest.log
tail -3 gpurun_out/y_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/y_smoke.log 2>&1; tail -1 gpurun_out/y_smoke.log
