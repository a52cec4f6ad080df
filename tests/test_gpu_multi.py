"""Multi-GPU parity on real devices (skipped on a single-GPU box): tools/mgpu_check.py under
torchrun with 2 ranks -- particle-decomposed sheath vs the same global state on one GPU."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_sheath_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "mgpu_check.py"), "200000"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [json.loads(l) for l in res.stdout.splitlines() if l.startswith("{")]
    out = [l for l in lines if "iters_sharded" in l][-1]
    assert out["ok"], out
    assert out["iters_sharded"] == out["iters_single"]
    per = [l for l in lines if "periodic" in l][-1]["periodic"]
    assert per["ok"], per
