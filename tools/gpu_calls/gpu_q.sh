#!/bin/bash
# round 2, call Q (8 GPUs): concurrent H2D ceiling at 1/2/4/8 ranks + the bench line at N = 8 (value + e2e, NUMA-pinned ranks)
cd /root/repo
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/q_topo.txt 2>&1
lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" > gpurun_out/q_lscpu.txt 2>&1
N=${NGPU:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/h2d_probe.py > gpurun_out/q_h2d_probe.json 2> gpurun_out/q_h2d_probe.err
tail -c 1500 gpurun_out/q_h2d_probe.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 100 --warmup 6 > gpurun_out/q_bench_n$N.json 2> gpurun_out/q_bench_n$N.err
head -c 400 gpurun_out/q_bench_n$N.json; echo
python - <<PY
import json
j=json.load(open("gpurun_out/q_bench_n$N.json"))
print("N=$N value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "gray", round(j["e2e_gray8"]["value"]), j["config"].get("host_affinity"))
PY
tail -3 gpurun_out/q_bench_n$N.err | cut -c1-300
