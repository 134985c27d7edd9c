"""Diagnostic (not a test): one window whose bounded landmarks sit on the bound, kernel-side timing of the line search."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vins-rgbd-fast_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ba_ref
from vrf_b200 import ba_problem as BP, binding as B
from test_ba_gpu import make_cfg
cfg = make_cfg()
h = B.Handle(cfg, 1, 0)
BP.WindowSimulator.FLAG2_DEPTH = (0.52, 1.2)
sim = BP.WindowSimulator(7, cfg, n_landmarks=80, preintegrate=ba_ref.preintegrate)
for a in range(2):
    pb = sim.window(a)
    so = ba_ref.solve(cfg, pb)
    if a == 1:
        sg = h.ba_solve(0, pb)
        print("gpu", sg.c.iterations, sg.c.successful_steps, sg.c.armijo_failures, "oracle", so.c.iterations, so.c.successful_steps, so.c.armijo_failures, flush=True)
    sim.commit(a, so)
