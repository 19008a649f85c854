"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): torchrun launches scripts/mgpu_check.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_rank_transform_matches_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "scripts", "mgpu_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert "MGPU_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
