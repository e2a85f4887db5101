"""Member-sharded ensembles on two GPUs of one box (skipped below two devices): tools/check_multi_gpu.py under
torchrun -- the sharded run, on the peer-read transform and on the all-gather path, equals the same ensemble
on one GPU."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def test_member_sharded_run_equals_single_gpu(libtxh):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "check_multi_gpu.py")]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    out = res.stdout + res.stderr
    assert res.returncode == 0, out[-3000:]
    lines = [l for l in out.splitlines() if l.startswith("multi-GPU check")]
    assert len(lines) == 2 and all(l.endswith("OK") for l in lines), out[-3000:]
