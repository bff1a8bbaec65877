"""Multi-GPU tests (need >= 2 GPUs: `gpurun --gpus 2`; skipped on a one-GPU box).  The multi-process form -- torchrun, one
rank per GPU, ONE NCCL communicator inside libgpann_cuda -- and the C++ drop-in over a group ctx ($GPANN_DEVICES)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


def test_sharded_entry_points_two_ranks():
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    world = min(_gpus(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "multi_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MULTI OK" in r.stdout, r.stdout[-3000:] + r.stderr[-6000:]


def test_dropin_over_a_group_context():
    """The C++ drop-in (ComputeGroundTruth, KMeans, ... with the reference's signatures) on 2 GPUs behind one call."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, GPANN_DEVICES="0,1")
    cmd = [sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", os.path.join(ROOT, "tests", "test_gpu_dropin.py"),
           os.path.join(ROOT, "tests", "test_integration.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1800, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
