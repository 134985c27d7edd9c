"""Bring-up diagnostic (not a test): per-iteration traces of oracle and GPU solvers."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vins-rgbd-fast_b200"))
import numpy as np
from oracle import ba_ref
from vrf_b200 import ba_problem as BP, binding as B
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_ba_gpu import make_cfg
cfg = make_cfg()
h = B.Handle(cfg, 1, 0)
from oracle import ba_ref as _br
sim = BP.WindowSimulator(5, cfg, n_landmarks=150, preintegrate=_br.preintegrate)
import ctypes as C
for a in range(2):
    pb = sim.window(a)
    if os.environ.get("NO_PRIOR") and a == 1:
        pb.prior = None; pb.finalize()
    os.environ["ORACLE_BA_DEBUG"] = "1"
    so = ba_ref.solve(cfg, pb)
    sys.stderr.flush()
    sg = h.ba_solve(0, pb)
    print("window", a, "oracle", so.c.iterations, so.c.successful_steps, so.c.final_cost, "gpu", sg.c.iterations, sg.c.successful_steps, sg.c.final_cost, flush=True)
    sim.commit(a, so)
