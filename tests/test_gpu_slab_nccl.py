"""The NCCL transport of the slab-sharded solves inside `-m gpu` (round-1 finding: it was only covered by a tool the driver
never runs).  Needs two GPUs: launches tests/_nccl_slab_worker.py with one process per GPU through torch.distributed.run and skips on a
one-GPU box (where the same slab code is covered with the in-process thread transport by tests/test_gpu_slab*.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_solves_over_nccl_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (one process per GPU over NCCL)")
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    port = 29600 + os.getpid() % 2000
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "_nccl_slab_worker.py")],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, (p.stdout[-3000:], p.stderr[-3000:])
    assert p.stdout.count("NCCL_SLAB_OK") == 2
