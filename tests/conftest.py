import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/_build/libbam3d_oracle.so), built on demand with gcc. Test infrastructure only."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    from tests import oracle_ffi
    return oracle_ffi.Oracle()


@pytest.fixture(scope="session")
def gpu_lib():
    """libbam3d_gpu.so, built on demand with nvcc (cross-compiles without a GPU)."""
    from bam3d_b200 import build
    build.build()
    from bam3d_b200 import _ffi
    return _ffi.lib()


@pytest.fixture(scope="session")
def ctx(gpu_lib):
    import bam3d_b200 as b3
    return b3.Context(0)
