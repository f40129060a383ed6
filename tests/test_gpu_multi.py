"""Multi-GPU path on real GPUs: one process per GPU under torchrun / NCCL.  Runs when the box has at least two
GPUs (skipped on a single-GPU box; the gloo world-2 tests in tests/test_distributed_gloo.py cover the host logic
everywhere).  The check itself is ``tools/dist_check.py``: distributed one-shot conditioning (block-row cyclic assembly +
Cholesky, replicated and non-replicated factor) == sequential single-GPU conditioning to 1e-9."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("npde,nbc", [(7000, 129), (3333, 64)])
def test_distributed_conditioning_matches_single_gpu(npde, nbc):
    import torch

    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if ngpu < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "dist_check.py"), str(npde), str(nbc)]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("posterior rel diff vs sequential single-GPU") == world
