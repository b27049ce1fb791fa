"""Multi-GPU paths on real devices (skipped on a one-GPU box; CPU-side logic is covered with gloo in test_cpu_host.py):
sharded sliding-window inference over 2 NCCL ranks must equal the single-rank result (tools/check_sw_sharded.py)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_sliding_window_equals_single_rank():
    n = 2
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                        os.path.join(ROOT, "tools", "check_sw_sharded.py")],
                       capture_output=True, text=True, cwd=ROOT, timeout=900)
    print(r.stdout[-4000:], r.stderr[-4000:])
    assert r.returncode == 0
    assert r.stdout.count("RESULT ok") == n
