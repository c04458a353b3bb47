"""The reference's own GJK tests, written against the C++ host layer (include/bam3d.hpp), compiled with g++ and run on
the GPU: checks that a native-language caller gets the crate's answers through the C ABI."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_kats_through_cpp_host_layer(gpu_lib, tmp_path):
    exe = str(tmp_path / "kat_gpu")
    libdir = os.path.join(ROOT, "bam3d_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", os.path.join(ROOT, "tests", "cpp", "kat_gpu.cpp"), "-o", exe,
                    "-L" + libdir, "-lbam3d_gpu", "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all passed" in r.stdout
