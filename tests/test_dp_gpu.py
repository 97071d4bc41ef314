"""GPU, 2 ranks: BASELINE.json configs[4] — the data-parallel step (NCCL gradient all-reduce inside the library) against the
single-device step on the concatenated batch, fp32 (2e-5) and bf16 (2e-2), replicas bit-identical.  Skips on a box with
fewer than two GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_dp_gpu.py -m gpu`)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.gpu
def test_dp_step_matches_single_device_step():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "scripts", "dp_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    assert "FAIL" not in r.stdout and "replicas bit-identical: True" in r.stdout
