"""N > 1 path on real GPUs: launches tests/multi_gpu_check.py under torchrun (needs >= 2 GPUs on the box)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
def test_sharded_runs_match_single_gpu():
    n = _gpu_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = 2
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(ROOT, "tests", "multi_gpu_check.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert f"multi_gpu_check ok: world={world}" in out.stdout
