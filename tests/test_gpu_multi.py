"""Multi-GPU parity (needs >= 2 GPUs on the box: `gpurun --gpus 2 -- python -m pytest tests -m gpu`).
Launches tests/mg_worker.py under torchrun, one rank per GPU; skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_downsample_and_kmeans(world):
    n = _n_gpus()
    if n < world:
        pytest.skip(f"needs {world} GPUs, box has {n}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world),
           os.path.join(ROOT, "tests", "mg_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    sys.stdout.write(r.stdout[-3000:])
    if r.returncode != 0:  # the workers' own tracebacks come before torchrun's failure summary
        at = r.stderr.find("Traceback")
        sys.stderr.write(r.stderr[max(0, at - 500):at + 6000] if at >= 0 else r.stderr[-6000:])
    else:
        sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0
    assert r.stdout.count("mg ok") == 10
