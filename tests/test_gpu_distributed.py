"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): slab-decomposed fftn with the
NCCL all-to-all and batch sharding, launched as one process per GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_slab_fftn_and_batch_sharding_two_gpus():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(here, "_slab_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_slab_fftn_rfftn_single_rank_group():
    """The same script on a one-rank process group (runs on a 1-GPU box): every exchange engine of SlabFFTN / SlabRFFTN --
    the fused transform + push kernel writing into its own symmetric buffer, the symmetric-memory push, NCCL
    all_to_all_single -- complex and real volumes, against the reference on the host."""
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "1", "--master-addr",
           "127.0.0.1", "--master-port", "29534", os.path.join(here, "_slab_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "engine fused" in r.stdout and "real" in r.stdout, r.stdout[-1500:]
