"""The boundary is a C ABI: the public headers must compile as plain C99, a C program must link against libvrf.so and
run without Python, and without a CUDA device the library must refuse to create a handle (no CPU fallback)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "vins-rgbd-fast_b200")


def test_headers_are_c99_and_a_c_program_links(built_lib, tmp_path):
    exe = str(tmp_path / "abi_smoke")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "abi_smoke.c"), "-o", exe, "-L", PKG, "-lvrf", "-Wl,-rpath," + PKG])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    txt = out.stdout
    assert "cfg 640x480 max_cnt 150" in txt
    assert "null handle rc -1 -1 -1" in txt
    import torch
    if not torch.cuda.is_available():
        assert "vrf_create rc -5" in txt and "handle null" in txt          # VRF_ERR_NO_DEVICE
    else:
        assert "vrf_create rc 0" in txt and "empty batches rc 0 0" in txt


def test_host_shims_instantiate_against_the_reference_member_names(built_lib, tmp_path):
    """vins-rgbd-fast_b200/host/*.h are templates over the reference's own types: instantiate them against stand-ins that
    expose exactly the members they touch (Eigen / OpenCV are not in the image), link against libvrf.so and run."""
    exe = str(tmp_path / "host_shims_check")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "host"),
                           os.path.join(ROOT, "tests", "host_shims_check.cpp"), "-o", exe, "-L", PKG, "-lvrf", "-Wl,-rpath," + PKG])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "shims instantiated: 4" in out.stdout
