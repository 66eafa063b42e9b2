"""GPU, >= 2 devices: the fused peer-store all-gather of the transforms (kernel epilogue writing into every rank's
buffer through NVLink peer pointers) must equal an NCCL all_gather, and the sharded hist_icp with the exchanged batch stop
must equal the unsharded call bit for bit.  Skipped on single-GPU boxes."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fused_peer_gather_matches_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29571", os.path.join(ROOT, "tools", "p2p_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("fused peer gather == NCCL all_gather: True") == 2
    assert out.stdout.count("peer push == NCCL all_gather: True") == 2
    assert out.stdout.count("sharded hist_icp == unsharded: True") == 2
