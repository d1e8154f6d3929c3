"""Multi-GPU correctness under pytest -m gpu: launches tests/multi_gpu_check.py on 2 ranks (NCCL, one process per
GPU) when the box has at least two B200s, in both exchange modes; skipped on a one-GPU box (there the in-run parity
probe of bench.py --gpus N is the driver-visible evidence)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.timeout(700)      # the subprocess has its own 600 s limit
@pytest.mark.parametrize("mode", ["p2p", "nccl"])
def test_two_rank_parity(mode):
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, APOP_B200_ALLREDUCE=mode)
    port = 29610 + (0 if mode == "p2p" else 1)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_check.py")],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
