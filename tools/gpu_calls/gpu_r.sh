#!/bin/bash
# round 2, call R (8 GPUs): the two multi-GPU BASELINE configurations as they are written --
#   configs[3]: 1280x720, 500 feats, 4 levels, 64 sequences over 4 GPUs;  configs[4]: 300 feats, full VIO, 256 sequences over 8 GPUs
cd /root/repo
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --config c4 --gpus 4 --steps 100 --warmup 6 > gpurun_out/r_bench_c4_n4.json 2> gpurun_out/r_bench_c4_n4.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --config c5 --gpus 8 --steps 100 --warmup 6 > gpurun_out/r_bench_c5_n8.json 2> gpurun_out/r_bench_c5_n8.err
python - <<PY
import json
for f in ("r_bench_c4_n4", "r_bench_c5_n8"):
    try:
        j=json.load(open("gpurun_out/%s.json" % f))
        print(f, "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "ms", round(j["ms_per_step"],3), j["config"]["seqs_per_gpu"], j["n_gpus"])
    except Exception as e:
        print(f, "failed", e)
PY
tail -2 gpurun_out/r_bench_c4_n4.err | cut -c1-300; tail -2 gpurun_out/r_bench_c5_n8.err | cut -c1-300
