"""N > 1 host logic on the CPU, `gloo`, world_size 2 and 3 (no GPU needed): the
worker is tests/gloo_check.py.  Also the row-block arithmetic of the C ABI."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_row_block_tiles_and_aligns():
    from lsbench_b200 import abi
    for n in (1, 31, 32, 33, 1000, 5570, 512 ** 3, 50_000_000):
        for P in (1, 2, 3, 4, 8):
            cuts = [abi.row_block(n, k, P) for k in range(P)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            for k in range(P - 1):
                assert cuts[k][1] == cuts[k + 1][0] and (cuts[k][1] % 32 == 0 or cuts[k][1] == n)
            assert all(0 <= a <= b <= n for a, b in cuts)
            if n >= 64 * P:
                sizes = [b - a for a, b in cuts]
                assert max(sizes) - min(sizes) <= 64
    # 512^3 on 8 ranks: whole z-planes (the halo is then two planes)
    assert abi.row_block(512 ** 3, 3, 8) == (3 * 64 * 512 * 512, 4 * 64 * 512 * 512)
    with pytest.raises(abi.B200Error):
        abi.row_block(100, 4, 4)


@pytest.mark.parametrize("nranks", [2, 3])
def test_row_block_path_emulated_over_gloo(nranks):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node", str(nranks), "--master-addr", "127.0.0.1",
           "--master-port", str(29510 + nranks), os.path.join(HERE, "gloo_check.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "GLOO_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
