"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the row-sharded paths -- NCCL gather and the
fused peer-memory gather -- against the oracle, one process per GPU under torchrun."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_sharded_paths_on_two_gpus():
    from sparse_matrix_mult_b200.matrix_ops import matrix_ops
    if matrix_ops.get_lib().spgemm_b200_device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "scripts", "multigpu_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "multigpu check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
