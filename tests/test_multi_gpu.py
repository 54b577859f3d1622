"""Multi-rank parity as a pytest: spawns tests/multi_gpu_parity.py under torch.distributed.run with 2 ranks (one GPU each)
when at least two GPUs are visible (skipped otherwise); the script compares vmult, the FDM preconditioner (symm / post / none /
ras) and a Chebyshev step on the partitioned mesh with the same mesh on one rank, 1e-12 relative (double)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4])
def test_multi_rank_parity(world):
    if _n_gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    port = 29700 + os.getpid() % 200 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTI_GPU_PARITY PASS" in r.stdout
