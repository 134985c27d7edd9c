"""Static properties of the shipped machine code (cuobjdump -sass on libvrf.so; no GPU needed): the instruction-level claims of
DESIGN.md section 3 as regression guards -- TMA loads and integer dot products in the front end, FP64 tensor-core MMA in the
Cholesky, and NO floating-point atomic loops anywhere in the back end (its sums have a fixed order: bit-reproducible results)."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "vins-rgbd-fast_b200", "libvrf.so")


@pytest.fixture(scope="module")
def sass_counts(built_lib):
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, timeout=600).stdout
    pats = {"UTMALDG": r"\bUTMALDG", "IDP.2A": r"\bIDP\.2A", "IDP.4A": r"\bIDP\.4A", "DMMA": r"\bDMMA", "ATOMS.CAS": r"\bATOMS\.CAS",
            "ATOMG": r"\bATOMG", "RED": r"\bRED\.", "ARCH": r"sm_100a|SM100"}
    cnt, cur = collections.defaultdict(collections.Counter), None
    archs = set(re.findall(r"arch = (sm_\w+)", txt))
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur:
            for k, p in pats.items():
                if re.search(p, line):
                    cnt[cur][k] += 1
    return cnt, archs


def _kernels(cnt, name):
    return {f: c for f, c in cnt.items() if name in f}


def test_library_is_built_for_sm_100a_only(sass_counts):
    _, archs = sass_counts
    assert archs == {"sm_100a"}, archs


def test_front_end_uses_tma_and_integer_dot_products(sass_counts):
    cnt, _ = sass_counts
    lk = _kernels(cnt, "k_lk")
    assert lk and all(c["UTMALDG"] >= 1 and c["IDP.2A"] >= 8 and c["IDP.4A"] >= 8 for c in lk.values()), lk
    pyr = _kernels(cnt, "k_pyr")
    assert len(pyr) == 3                                           # RGB8 frame, GRAY8 frame, pyramid level
    assert all(c["IDP.4A"] >= 16 for c in pyr.values()), pyr
    assert sum(c["IDP.2A"] >= 32 for c in pyr.values()) == 1        # the RGB -> gray conversion of the RGB8 variant


def test_cholesky_runs_on_the_fp64_tensor_cores(sass_counts):
    cnt, _ = sass_counts
    solve = _kernels(cnt, "k_ba_solve")
    assert solve and all(c["DMMA"] >= 2 for c in solve.values()), solve


def test_back_end_has_no_floating_point_atomics(sass_counts):
    """A floating-point atomicAdd compiles to an ATOMS.CAS loop (shared memory) or RED / ATOMG (global memory).  None may exist in
    the bundle-adjustment, marginalization, prior-factor, feature-manager or pre-integration kernels."""
    cnt, _ = sass_counts
    for name in ("k_ba_solve", "k_ba_marg", "k_ba_prior_factor", "k_fm_", "k_imu_preint"):
        ks = _kernels(cnt, name)
        assert ks, name
        for f, c in ks.items():
            assert c["ATOMS.CAS"] == 0 and c["ATOMG"] == 0 and c["RED"] == 0, (f, dict(c))
