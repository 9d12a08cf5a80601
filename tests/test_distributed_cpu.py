"""World-size-2 and -3 tests of the multi-GPU HOST logic on CPU (gloo): the tet partition every rank derives
independently must be consistent across ranks, and the exchange it implies -- each rank sums the dx of
its own tets, the shared-boundary sums are all-reduced, every rank applies the reduced value -- must
reproduce the single-process Jacobi iteration.  The per-tet arithmetic is the oracle's (this is test
infrastructure on CPU; the GPU version of the same check is tools/multigpu_check.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, cluster_size, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from tetsim_b200 import _capi, mesh
    v, t = mesh.make_beam((24, 6, 5), h=0.05, y0=0.5, jitter=0.2)
    N, M = v.size // 3, t.size // 4
    plan = _capi.plan_partition(v, t, cluster_size, True, rank, world)
    l2c, nI, nB = plan["localToCaller"], plan["numInterior"], plan["numBoundary"]
    # ---- consistency of independently derived plans ----
    counts = torch.tensor([len(plan["localTets"]), nI, nB], dtype=torch.int64)
    allc = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(allc, counts)
    assert sum(int(c[0]) for c in allc) == M, "tets are not partitioned exactly"
    assert len({int(c[2]) for c in allc}) == 1, "ranks disagree on the boundary set size"
    bset = torch.from_numpy(l2c[nI:].astype(np.int64))
    allb = [torch.zeros_like(bset) for _ in range(world)]
    dist.all_gather(allb, bset)
    assert all(torch.equal(allb[0], b) for b in allb), "boundary sets differ or are ordered differently"
    own = torch.zeros(M, dtype=torch.int32)
    own[torch.from_numpy(plan["localTets"].astype(np.int64))] = 1
    dist.all_reduce(own)
    assert int(own.min()) == 1 and int(own.max()) == 1, "a tet is owned by zero or two ranks"
    tets = t.reshape(-1, 4)
    touched = np.unique(tets[plan["localTets"]])
    assert set(touched) <= set(l2c.tolist()), "a local tet references a non-resident vertex"
    interior = torch.zeros(N, dtype=torch.int32)
    interior[torch.from_numpy(l2c[:nI].astype(np.int64))] = 1
    dist.all_reduce(interior)
    assert int(interior.max()) <= 1, "an interior vertex is resident on two ranks"
    # ---- the exchange: local sums + all-reduce of the boundary == the single-process iteration ----
    dt = 1.0 / 1200.0
    body = oracle.SoftBodyOracle(v, t)
    body.pos[1::3] -= np.float32(0.01) * np.arange(N, dtype=np.float32) / N   # some deformation
    acc = body.jacobi_accumulate(plan["localTets"], dt).reshape(N, 3)
    b = torch.from_numpy(acc[l2c[nI:]].copy())
    dist.all_reduce(b)                                     # what ncclAllReduce does on the GPU path
    merged = np.full((N, 3), np.nan, np.float32)
    merged[l2c[:nI]] = acc[l2c[:nI]]
    merged[l2c[nI:]] = b.numpy()
    full = body.jacobi_accumulate(np.arange(M, dtype=np.int32), dt).reshape(N, 3)
    res = ~np.isnan(merged[:, 0])
    scale = np.abs(full).max()
    assert np.max(np.abs(merged[res] - full[res])) <= 2e-6 * scale
    # ---- the neighbour / peer-memory exchanges: every sharer adds the ranks' sums in ascending rank order ----
    mine = torch.from_numpy(acc[l2c[nI:]].copy())          # this rank's sums of the boundary vertices (zeros where it has no tet)
    every = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(every, mine)
    ordered = np.zeros((nB, 3), np.float32)
    for q in range(world):                                 # float32, rank 0 first: the order k_halo_reduce / k_jacobi_apply<PEER> use
        ordered = ordered + every[q].numpy()
    mine_bits = torch.from_numpy(ordered.view(np.int32).copy())
    all_bits = [torch.zeros_like(mine_bits) for _ in range(world)]
    dist.all_gather(all_bits, mine_bits)
    assert all(torch.equal(all_bits[0], x) for x in all_bits), "rank-ordered sums must be bit-identical on every rank"
    assert np.max(np.abs(ordered - full[l2c[nI:]])) <= 2e-6 * scale
    # every vertex is resident somewhere
    r = torch.from_numpy(res.astype(np.int32))
    dist.all_reduce(r)
    assert int(r.min()) >= 1
    out.put((rank, nI, nB))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,cluster_size", [(2, 64), (2, 256), (3, 128)])
def test_partition_and_exchange(world, cluster_size):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cluster_size, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    got = sorted(q.get(timeout=5) for _ in range(world))
    assert len({g[2] for g in got}) == 1 and got[0][2] > 0


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_peer_exchange_layouts_agree_across_ranks(world):
    """The peer-memory exchange communicates nothing but memory handles: every rank DERIVES where its entries land
    in each sharer's receive buffer.  Check those derived layouts against the sharers' own (host only, no GPU)."""
    from tetsim_b200 import _capi, mesh
    v, t = mesh.make_beam((40, 6, 6), h=0.05, y0=0.5, jitter=0.2)
    plans = [_capi.plan_halo(v, t, 128, True, r, world) for r in range(world)]
    pairs = 0
    for r, p in enumerate(plans):
        assert list(p["peers"]) == sorted(set(p["peers"])) and r not in p["peers"]
        for i, q in enumerate(p["peers"]):
            pq = plans[q]
            assert r in pq["peers"], "sharing must be symmetric"
            j = list(pq["peers"]).index(r)
            seg_r = p["segStart"][i + 1] - p["segStart"][i]
            seg_q = pq["segStart"][j + 1] - pq["segStart"][j]
            assert seg_r == seg_q > 0, "both sides exchange the same shared set"
            assert p["remoteOff"][i] == pq["segStart"][j], "my entries start where the sharer expects them"
            assert p["remoteTotal"][i] == pq["segStart"][-1], "parity stride of the sharer's buffer"
            assert p["remoteSlot"][i] == j, "which of the sharer's flags is mine"
            pairs += 1
    assert pairs >= 2 * (world - 1)


# ------------------------------------------------------------------------------------------------
# BASELINE config 5 on N GPUs: independent bodies sharded across ranks, no exchange (tetsim_b200.mesh.shard_bodies)
# ------------------------------------------------------------------------------------------------
def _shard_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from tetsim_b200 import mesh
    m = mesh.load_dragon()
    v, t = mesh.tile_bodies(m["tet_verts"], m["tet_ids"], 3, 2, y_shift=-0.40)     # 6 bodies
    N, M, nb = v.size // 3, t.size // 4, 6
    sv, st, vid, tix = mesh.shard_bodies(v, t, rank, world)
    # every vertex and tet on exactly one rank; bodies whole; the reference's sweep order kept inside every body
    own_v = torch.zeros(N, dtype=torch.int32); own_v[torch.from_numpy(vid.astype(np.int64))] = 1
    own_t = torch.zeros(M, dtype=torch.int32); own_t[torch.from_numpy(tix.astype(np.int64))] = 1
    dist.all_reduce(own_v); dist.all_reduce(own_t)
    assert int(own_v.min()) == 1 and int(own_v.max()) == 1 and int(own_t.min()) == 1 and int(own_t.max()) == 1
    assert np.all(np.diff(tix) > 0) and np.all(np.diff(vid) > 0)
    assert (sv.size // 3) % 1234 == 0 and (st.size // 4) % 3840 == 0
    counts = torch.tensor([st.size // 4 // 3840], dtype=torch.int64)
    allc = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(allc, counts)
    assert sum(int(c) for c in allc) == nb and max(int(c) for c in allc) - min(int(c) for c in allc) <= 1
    assert np.array_equal(sv.reshape(-1, 3), v.reshape(-1, 3)[vid])
    assert np.array_equal(vid[st.reshape(-1, 4)], t.reshape(-1, 4)[tix])
    # a shard simulated alone == those bodies inside the whole scene, bit for bit (per-copy check on the SAME translated copy)
    wb = mesh.wide_bounds(64.0)
    whole = oracle.SoftBodyOracle(v, t, worldBounds=wb)
    part = oracle.SoftBodyOracle(sv, st, worldBounds=wb)
    for _ in range(6):
        whole.simulate(1.0 / 600.0)
        part.simulate(1.0 / 600.0)
    assert np.array_equal(part.pos.reshape(-1, 3).view(np.uint32), whole.pos.reshape(-1, 3)[vid].view(np.uint32))
    merged = torch.zeros(3 * N, dtype=torch.float32)
    merged.view(-1, 3)[torch.from_numpy(vid.astype(np.int64))] = torch.from_numpy(part.pos.copy()).view(-1, 3)
    dist.all_reduce(merged)       # disjoint supports: a sum with zeros is exact
    assert np.array_equal(merged.numpy().view(np.uint32), whole.pos.view(np.uint32))
    out.put(rank)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_body_sharding_no_exchange(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert sorted(out.get(timeout=5) for _ in range(world)) == list(range(world))
