"""The library's NCCL contact exchange across processes, through the C ABI: tests/tools/gather_2gpu.py under torchrun, one rank per
GPU (every rank runs GJK+EPA on its slice, then exchanges through the device-pointer calls with two steps in flight, the
host-buffer call and the frame queue's gather mode, and compares with the single-process hit list of the whole scene).
Needs two GPUs: skipped on the single-GPU box the driver's `pytest -m gpu` runs on; run by hand on 2 / 8 GPUs this round."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2])
def test_contact_exchange_across_ranks_through_the_c_abi(gpu_lib, world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "tools", "gather_2gpu.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert f"gather ok on {world} ranks" in r.stdout
