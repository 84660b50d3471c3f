"""GPU (needs >= 2 devices): a 2-rank NCCL data-parallel run (gradient buckets + SyncBN) on a global batch equals
the single-process run on the same batch - parameters, BN running statistics and losses after two SGD steps."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_data_parallel_equals_single_process():
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "scripts", "dp_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, "DP_MATH": "fp32"})
    assert "DP_PARITY_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
