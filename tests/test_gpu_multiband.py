"""Split frame on real GPUs: two bands (one process per GPU, torchrun) exchange their strips inside the library -- peer-memory
stores from the frame kernel and the NCCL all-gather -- and every band's frame must equal the single-band render.
Skipped on a one-GPU box (the gloo tests in test_split_frame_gloo.py cover the host logic there)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_band_exchange_matches_single_band():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "check_band_gather.py"), "both"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "BAND EXCHANGE OK" in r.stdout
