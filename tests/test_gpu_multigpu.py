"""NCCL path of libsbd (sbd_comm_init + the per-iteration all-gather) under pytest: chains sharded over 2 ranks must
give trajectories bit-identical on every rank AND to one rank running all the chains - also at a size (1024^2) where
the launch geometry used to depend on the number of chains per rank.  Needs >= 2 GPUs; skipped otherwise
(the 1-GPU box of the round-end run); run with `gpurun --gpus 2`."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("size,per,samples", [(128, 2, 30), (1024, 2, 10), (1024, 4, 8)])
def test_sharded_chains_are_bitwise_identical(size, per, samples):
    if ngpus() < 2:
        pytest.skip("needs >= 2 GPUs")
    port = 29530 + (size // 128) % 50 + per
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tools", "check_multigpu.py"), "--size", str(size), "--per", str(per), "--samples", str(samples)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ranks bit-identical: True" in r.stdout
    assert "multi-GPU == single-GPU (bitwise): True" in r.stdout
    assert "CHECK PASS" in r.stdout
