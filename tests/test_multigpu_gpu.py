"""Multi-GPU Jacobi (tet partition + ncclAllReduce of boundary dx): runs tools/multigpu_check.py under
torchrun when the box has at least two B200s, otherwise skips (the driver's -m gpu box has one)."""
import os
import subprocess
import sys

import pytest

from tetsim_b200 import _capi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world,exchange", [(2, "allreduce"), (2, "halo"), (4, "allreduce"), (4, "halo")])
def test_partitioned_jacobi_matches_single_gpu(world, exchange):
    n = _capi.lib().tetsim_device_count()
    if n < world:
        pytest.skip("needs %d GPUs, box has %d" % (world, n))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29533 + world),
           os.path.join(ROOT, "tools", "multigpu_check.py"), "--cells", "64,16,16", "--substeps", "40", "--exchange", exchange]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
