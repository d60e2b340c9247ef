"""Multi-GPU parity of the sharded K4 path (needs >= 2 B200s in the process' view; skipped on a single-GPU box, where
the exchange logic is covered by the world-2 gloo tests and the merge kernels by tests/test_gpu_topk.py)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_sharded_topk_over_nccl_equals_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multi_gpu_check.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), script], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "multi-gpu ok" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
