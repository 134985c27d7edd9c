"""Not a test: diagnostic dump used during bring-up (python tests/debug_front.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vins-rgbd-fast_b200"))
import numpy as np, cv2
from oracle.frontend_ref import FeatureTrackerRef, FrontendConfig
from oracle import frontend_spec as S
from vrf_b200 import binding, synth

seq = synth.Sequence(1234)
cfg = binding.default_config(use_ransac=0)
h = binding.Handle(cfg, 1, 0)
ref = FeatureTrackerRef(FrontendConfig(use_ransac=0))
for k in range(7):
    _, gray, _ = seq.frame(k)
    R = seq.relative_R(k); pub = (k % 3 == 0)
    out = h.read_image(0, gray, seq.time(k), R, pub=pub)
    ref.read_image(gray, seq.time(k), R, pub_this_frame=pub)
    for l in range(2):
        pl = h.pyramid_level(0, l)
        cvl = gray if l == 0 else cv2.pyrDown(gray)
        print("frame", k, "pyr", l, "equal", np.array_equal(pl, cvl))
    print("frame", k, "n", out.n, len(ref.ids))
    if pub:
        ck = h.debug_read("cell_k", 0, np.int32, 56)
        nc = h.debug_read("ncand", 0, np.int32, 56)
        cand = h.debug_read("cand", 0, np.float32, 56 * 4 * 3).reshape(56, 4, 3)
        gi = 0
        for cell in range(56):
            if cell in ref.last_grids_id:
                rk = ref.last_new_keypoints[ref.last_grids_id.index(cell)]
                mine = [tuple(cand[cell, j]) for j in range(nc[cell])]
                theirs = [(float(a), float(b), float(c)) for (a, b, c) in rk]
                if mine != theirs:
                    print(" cell", cell, "K", ck[cell], "MISMATCH\n   gpu", mine, "\n   ref", theirs)
            elif ck[cell] != 0:
                print(" cell", cell, "selected on gpu only")
    if ref.last_lk_pts is not None:
        if len(out.lk_status) == len(ref.last_lk_status):
            st_eq = np.array_equal(out.lk_status, ref.last_lk_status)
            ok = ref.last_lk_status.astype(bool) & out.lk_status.astype(bool)
            dd = np.abs(out.lk_pts[ok] - ref.last_lk_pts[ok])
            print("  lk status eq", st_eq, "n", len(ok), "ok", ok.sum(), "maxdiff", dd.max(), "p99", np.percentile(dd, 99),
                  "pred maxdiff", np.abs(out.predict_pts - ref.predict_pts).max())
            if not st_eq:
                bad = np.nonzero(out.lk_status != ref.last_lk_status)[0]
                print("   status mismatch at", bad, out.lk_pts[bad], ref.last_lk_pts[bad], ref.cur_pts[:0])
        else:
            print("  lk count differs", len(out.lk_status), len(ref.last_lk_status))
    m = min(out.n, len(ref.ids))
    print("  ids eq", np.array_equal(out.ids[:m], np.array(ref.ids[:m])), "pts maxdiff", np.abs(out.cur_pts[:m] - ref.cur_pts[:m]).max() if m else None)
