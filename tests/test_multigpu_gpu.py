"""Multi-GPU Jacobi (tet partition + ncclAllReduce of boundary dx): runs tools/multigpu_check.py under
torchrun when the box has at least two B200s, otherwise skips (the driver's -m gpu box has one)."""
import os
import subprocess
import sys

import pytest

from tetsim_b200 import _capi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world,exchange", [(2, "allreduce"), (2, "halo"), (2, "peer"), (4, "allreduce"), (4, "halo"),
                                            (4, "peer")])
def test_partitioned_jacobi_matches_single_gpu(world, exchange):
    n = _capi.lib().tetsim_device_count()
    if n < world:
        pytest.skip("needs %d GPUs, box has %d" % (world, n))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29533 + world),
           os.path.join(ROOT, "tools", "multigpu_check.py"), "--cells", "64,16,16", "--substeps", "40", "--exchange", exchange]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("world,deterministic,fused", [(2, True, True), (3, True, True), (4, True, True), (2, True, False),
                                                       (2, False, False)])
def test_peer_exchange_single_process(world, deterministic, fused, monkeypatch):
    """The peer-memory exchange protocol (push into the sharers' buffers + epoch flags, wait + rank-ordered reduce)
    driven on ONE GPU: `world` handles of this process, each owning one tet partition, are each other's peers (the
    blob carries the owner's pointer, so no cudaIpc mapping is involved).  Merged positions must agree with the
    unpartitioned body to 1e-5 and replicas of shared vertices must be bit-identical on every rank."""
    import numpy as np
    import tetsim_b200 as ts
    from tetsim_b200 import mesh

    monkeypatch.setenv("TETSIM_PEER_TIMEOUT_MS", "3000")   # a broken protocol fails in seconds instead of stalling the box
    # fused: the tile kernel pushes and the vertex kernel waits + reduces (2 launches per iteration, the default with the
    # deterministic flush); unfused: boundary tiles / push / interior tiles / wait + reduce / vertex kernel
    monkeypatch.setenv("TETSIM_PEER_UNFUSED", "0" if fused else "1")
    v, t = mesh.make_beam((48, 10, 10), h=0.02, y0=0.004, jitter=0.15)   # reaches the floor within the 60 substeps
    N = v.size // 3
    pp = dict(ts.DEFAULT_PHYSICS_PARAMS, numSubsteps=20, worldBounds=list(mesh.wide_bounds(16.0)))
    kw = dict(solver="jacobi", iters=2, cluster_size=128, deterministic=deterministic)
    bodies = [ts.SoftBody(v, t, None, pp, rank=r, world_size=world, exchange="peer", **kw) for r in range(world)]
    assert all(b.info()["launchesPerSubstep"] == (2 * 2 if fused else 4 * 2) for b in bodies)
    with pytest.raises(ts.TetSimError):
        bodies[0].step(pp)                       # peers not set yet
    blobs = [b.ipc_handle() for b in bodies]
    for b in bodies:
        b.set_peers(blobs)
    for _ in range(3):                           # launches are asynchronous: rank r's wait is satisfied by the
        for b in bodies:                         # pushes of the ranks enqueued after it
            b.step(pp)
    for b in bodies:
        b.synchronize()
    single = ts.SoftBody(v, t, None, pp, **kw)
    for _ in range(3):
        single.step(pp)
    ref = single.pos.reshape(N, 3).astype(np.float64)
    P = np.stack([b.pos.reshape(N, 3) for b in bodies])
    R = np.stack([b.resident.astype(bool) for b in bodies])
    assert R.any(axis=0).all()
    merged = np.zeros((N, 3), np.float32)
    for r in range(world):
        merged[R[r]] = P[r][R[r]]
    shared = R.sum(axis=0) > 1
    assert shared.sum() > 0
    if deterministic:
        for r in range(world):
            sel = R[r] & shared
            assert np.array_equal(P[r][sel].view(np.uint32), merged[sel].view(np.uint32)), "replicas differ on rank %d" % r
    err = float(np.max(np.linalg.norm(merged - ref, axis=1) / np.linalg.norm(ref, axis=1)))
    assert err <= 1e-5, err
    assert np.any(ref[:, 1] == 0.0), "the scene is meant to reach the floor"
    for b in bodies:
        b.close()


def test_config5_body_sharding_bitexact_per_copy():
    """BASELINE config 5 across ranks: the tiled-Dragon scene sharded by bodies (no exchange).  Driven on ONE GPU with one
    handle per rank: every shard must equal the oracle run on the SAME translated copies bit for bit (reference
    arithmetic), and the merged shards must equal the unsharded scene on the GPU."""
    import numpy as np
    import oracle
    import tetsim_b200 as ts
    from tetsim_b200 import mesh

    m = mesh.load_dragon()
    v, t = mesh.tile_bodies(m["tet_verts"], m["tet_ids"], 3, 2, y_shift=-0.45)   # floor contact within the run
    wb = list(mesh.wide_bounds(64.0))
    pp = dict(ts.DEFAULT_PHYSICS_PARAMS, numSubsteps=10, worldBounds=wb)
    world = 3
    merged = np.full(v.size, np.nan, np.float32)
    for r in range(world):
        sv, st, vid, _ = mesh.shard_bodies(v, t, r, world)
        sb = ts.SoftBody(sv, st, None, pp, solver="gs_exact", arithmetic="bitexact")
        assert sb.info()["numComponents"] == 2 and sb.info()["bodyKernel"] == 1
        ref = oracle.SoftBodyOracle(sv, st, worldBounds=wb)
        for _ in range(3):
            sb.step(pp)
            for _ in range(10):
                ref.simulate((1.0 / 60.0) / 10)
        assert np.array_equal(sb.pos.view(np.uint32), ref.pos.view(np.uint32)), "shard %d differs from the oracle on the same copies" % r
        merged.reshape(-1, 3)[vid] = sb.pos.reshape(-1, 3)
        sb.close()
    whole = ts.SoftBody(v, t, None, pp, solver="gs_exact", arithmetic="bitexact")
    for _ in range(3):
        whole.step(pp)
    assert np.array_equal(merged.view(np.uint32), whole.pos.view(np.uint32))
    assert np.any(whole.pos.reshape(-1, 3)[:, 1] == 0.0), "the scene is meant to reach the floor"
