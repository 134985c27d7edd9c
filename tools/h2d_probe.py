#!/usr/bin/env python
"""Concurrent host->device copy ceiling of one box: every rank copies the e2e arm's per-step payload
(444 separately addressed 921.6 KB RGB8 frames from pinned host memory, four copy streams) to its own GPU, first one rank alone,
then 2, 4, 8 ranks at the same time.  Launch like the bench:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/h2d_probe.py
Rank 0 prints one JSON line: per-rank and aggregate GB/s at each concurrency, with and without NUMA pinning of rank + buffers."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    import bench
    out = {"world": world}
    for pinned in (False, True):
        aff = bench.pin_to_gpu_numa_node(lr) if pinned else "unpinned"
        n, fb = 444, 640 * 480 * 3
        hs = [torch.empty(fb, dtype=torch.uint8).pin_memory() for _ in range(96)]       # 96 distinct host frames, like the bench
        for h in hs:
            h.random_(0, 255)
        d = torch.empty(n * fb, dtype=torch.uint8, device="cuda")
        ss = [torch.cuda.Stream() for _ in range(4)]
        res = {}
        conc = 1
        while conc <= world:
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            gbs = 0.0
            if rank < conc:
                t = time.perf_counter()
                for rep in range(10):
                    for i in range(n):
                        with torch.cuda.stream(ss[i % 4]):
                            d[i * fb:(i + 1) * fb].copy_(hs[i % 96], non_blocking=True)
                torch.cuda.synchronize()
                gbs = 10 * n * fb / (time.perf_counter() - t) / 1e9
            v = torch.tensor([gbs], dtype=torch.float64, device="cuda")
            lst = [torch.zeros_like(v) for _ in range(world)]
            if world > 1:
                dist.all_gather(lst, v)
            else:
                lst = [v]
            per = [float(x.item()) for x in lst][:conc]
            res[str(conc)] = {"per_rank_gbs": [round(x, 1) for x in per], "aggregate_gbs": round(sum(per), 1), "min_gbs": round(min(per), 1)}
            conc *= 2
        out["numa_pinned" if pinned else "unpinned"] = {"affinity_rank0": aff, "concurrency": res}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
