"""The library's own NCCL gather on more than one GPU: skipped on a one-GPU box, run with
`gpurun --gpus 2`; the host-side partition logic is covered on CPU by tests/test_sharding_gloo.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_library_gather_matches_torch_all_gather():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    n = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(ROOT, "tools", "nccl_gather_check.py")]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "'all_ranks_ok': True" in res.stdout
