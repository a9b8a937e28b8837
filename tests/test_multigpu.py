"""Multi-GPU parity (needs >= 2 GPUs: run on the box with `gpurun --gpus 2 -- python -m pytest tests -m gpu`)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_node_partitioned_solves_match_golden(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and f"MGPU_OK world={world}" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
