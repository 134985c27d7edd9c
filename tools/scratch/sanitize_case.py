"""compute-sanitizer target: one BA window chain (solve + marginalization, bounded landmarks next to their bound so that the
line search runs) and one fused ingest + pyramid call, through the C ABI."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vins-rgbd-fast_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle import ba_ref
from vrf_b200 import ba_problem as BP, binding as B
from test_ba_gpu import make_cfg
BP.WindowSimulator.FLAG2_DEPTH = (0.52, 1.2)
which = sys.argv[1] if len(sys.argv) > 1 else "ba"
cfg = make_cfg()
if which == "ba":
    h = B.Handle(cfg, 1, 0)
    sim = BP.WindowSimulator(7, cfg, n_landmarks=60, preintegrate=ba_ref.preintegrate)
    for a in range(2):
        pb = sim.window(a)
        so = ba_ref.solve(cfg, pb)
        sg = h.ba_solve(0, pb)
        print("window", a, "gpu", sg.c.iterations, sg.c.successful_steps, sg.c.armijo_failures, "oracle", so.c.iterations, so.c.successful_steps, so.c.armijo_failures, flush=True)
        sim.commit(a, so)
    h.close()
elif which == "ba_ex":
    cfg = make_cfg(estimate_td=1)
    h = B.Handle(cfg, 1, 0)
    sim = BP.WindowSimulator(33, cfg, n_landmarks=50, td_true=0.02, td_constant=0, ex_constant=0, ex_perturb=0.02, preintegrate=ba_ref.preintegrate)
    pb = sim.window(0)
    sg = h.ba_solve(0, pb)
    print("ex/td window gpu", sg.c.iterations, sg.c.successful_steps, flush=True)
    h.close()
else:
    rng = np.random.default_rng(1)
    c2 = B.default_config(row=250, col=336, fx=300.0, fy=300.0, cx=168.0, cy=125.0, use_ransac=1, lk_max_level=2, max_cnt=60, min_dist=12)
    h = B.Handle(c2, 2, 0)
    for k in range(2):
        frames = [rng.integers(0, 256, (250, 336, 3), dtype=np.uint8) for _ in range(2)]
        h.read_image_batch([0, 1], frames, [0.1 * k, 0.1 * k], pubs=[1, 1])
    print("front ok", flush=True)
    h.close()
