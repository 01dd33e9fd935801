"""CPU test of the N > 1 host logic: world_size 2 over gloo (no GPU).  The ownership rules the CUDA
library implements on the device are stated in <package>/sharding.py; tests/gloo_worker.py runs
the whole exchange protocol over them with the oracle as the per-rank worker and compares with the
single-process oracle."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharding_protocol_world2_gloo():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29731",
           os.path.join(ROOT, "tests", "gloo_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    sys.stdout.write(r.stdout[-2000:])
    sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0
    assert "gloo ok" in r.stdout
