"""Data parallelism on real GPUs (needs >= 2 devices; skipped on a one-GPU box): N-rank averaged gradient == mean of
the single-rank shard gradients with BatchNorm local (reference srgan/trainer.py:143-157)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_gradient_equals_mean_of_shard_gradients():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tools", "dp_check.py"), "16"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert lines, out.stdout[-2000:] + out.stderr[-4000:]
    r = json.loads(lines[-1])
    assert r["broadcast_max_abs"] == 0.0, r          # attach(): rank 0's parameters everywhere
    assert r["dp_parity"] <= 1e-5, r                 # exchange == mean of per-shard gradients
    assert r["weight_drift_max_abs"] == 0.0, r       # identical updates keep the replicas identical
    assert all(abs(x) < 1e3 for x in r["losses"]), r
