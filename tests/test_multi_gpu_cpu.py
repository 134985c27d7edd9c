"""Host logic of the N>1 path (bench.py): sequence sharding and the max-over-ranks reduction of the
timed interval, exercised with world_size 2 on the gloo backend (no GPU needed).  The data path has no
collective: each rank owns its own sequences (weak scaling)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import bench


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    S, steps = 12, 5
    seqs = bench.shard_sequences(S, rank, world)
    ms_local = 10.0 * (rank + 1)            # pretend rank 1 is slower
    ms, frames = bench.aggregate_timing(ms_local, len(seqs) * steps, dist)
    out[rank] = (seqs, ms, frames)
    dist.destroy_process_group()


def test_two_rank_sharding_and_aggregation():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    s0, ms0, f0 = out[0]
    s1, ms1, f1 = out[1]
    assert len(s0) == len(s1) == 12                     # weak scaling: per-rank work fixed
    assert set(s0).isdisjoint(s1) and sorted(s0 + s1) == list(range(24))
    assert ms0 == ms1 == 20.0                           # max over ranks
    assert f0 == f1 == 2 * 12 * 5                       # whole-job frames


def test_pingpong_plan_is_continuous():
    """Ping-pong playback: consecutive steps use adjacent frames and inverse rotations going backwards."""
    T = 6
    idx_prev = None
    for step in range(40):
        idx, pidx = bench.frame_plan(step, T)
        assert 0 <= idx < T
        if step > 0:
            assert abs(idx - idx_prev) == 1 and pidx == idx_prev
        idx_prev = idx
    R = np.stack([np.eye(3)] + [np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]])] * (T - 1))
    assert np.allclose(bench.rel_rotation(R, 3, 2), R[3])
    assert np.allclose(bench.rel_rotation(R, 2, 3), R[3].T)


def test_gpu_arm_never_touches_the_oracle():
    """oracle/ is test infrastructure: bench.py's GPU arm (main) must not import it -- only the CPU legs
    (run_reference / _ref_worker, executed in child processes) may -- and the product package never imports it."""
    import inspect
    import bench
    import re
    src = inspect.getsource(bench.main)
    assert not re.search(r"(from|import)\s+oracle", src)
    for fn in (bench.run_reference, bench._ref_worker):
        assert fn.__module__ == "bench"
    import vrf_b200.binding as binding
    assert not re.search(r"(from|import)\s+oracle", inspect.getsource(binding))
    import vrf_b200.ba_problem as ba_problem
    import vrf_b200.synth as synth_mod
    for mod in (ba_problem, synth_mod):
        assert not re.search(r"(from|import)\s+oracle", inspect.getsource(mod))
