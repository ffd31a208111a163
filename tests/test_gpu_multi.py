"""Row-block multi-GPU parity (needs >= 2 GPUs; skipped on a 1-GPU box).
The worker is tests/dist_check.py, one rank per GPU under torchrun."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpus():
    from lsbench_b200 import abi
    try:
        return abi.device_count()
    except abi.B200Error:
        return 0


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_row_block_partition_matches_oracle(nranks):
    if _ngpus() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node", str(nranks), "--master-addr", "127.0.0.1",
           "--master-port", str(29400 + nranks), os.path.join(HERE, "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "DIST_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
