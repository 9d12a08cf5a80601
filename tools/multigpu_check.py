#!/usr/bin/env python3
"""Multi-GPU Jacobi check, launched with torchrun (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/multigpu_check.py [--cells 64,16,16] [--substeps 40]

Every rank builds the same beam, takes its tet partition (tetsim_create with rank/worldSize), and
steps it with an ncclAllReduce of the shared-boundary dx each Jacobi iteration.  Rank 0 also runs the
whole mesh on its own GPU; the merged N-rank positions must agree with it to 1e-5 (vector-relative;
the only difference is the summation order of the boundary partial sums), and replicas of a boundary
vertex must be BIT-identical on every rank that holds it.
"""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tetsim_b200 as ts  # noqa: E402
from tetsim_b200 import _capi, mesh  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", default="64,16,16")
ap.add_argument("--substeps", type=int, default=40)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--cluster-size", type=int, default=256)
ap.add_argument("--exchange", default="allreduce", choices=["allreduce", "halo", "peer"])
a = ap.parse_args()

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cells = tuple(int(c) for c in a.cells.split(","))
v, t = mesh.make_beam(cells, h=0.02, y0=0.3, jitter=0.15)
N = v.size // 3

nccl_id = None
if a.exchange != "peer":   # the peer-memory exchange needs no NCCL communicator of its own
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        raw = ctypes.create_string_buffer(128)
        _capi.check(_capi.lib().tetsim_nccl_unique_id(raw))
        buf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    nccl_id = bytes(buf.cpu().numpy().tobytes())

pp = dict(ts.DEFAULT_PHYSICS_PARAMS, numSubsteps=20, worldBounds=list(mesh.wide_bounds(16.0)))
body = ts.SoftBody(v, t, None, pp, solver="jacobi", iters=a.iters, cluster_size=a.cluster_size, device=local,
                   rank=rank, world_size=world, nccl_unique_id=nccl_id, exchange=a.exchange)
if a.exchange == "peer":   # hand-shake: all-gather the exchange-buffer handles, map the sharers'
    mine = torch.frombuffer(bytearray(body.ipc_handle()), dtype=torch.uint8).cuda()
    blobs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(blobs, mine)
    body.set_peers([bytes(b.cpu().numpy().tobytes()) for b in blobs])
    dist.barrier()
info = body.info()
for _ in range(a.substeps // 20):
    body.step(pp)
pos = torch.from_numpy(body.pos.copy()).cuda()          # NaN where not resident
vel = torch.from_numpy(body.vel.copy()).cuda()
res = torch.from_numpy(body.resident.astype(np.uint8)).cuda()

allpos = [torch.empty_like(pos) for _ in range(world)]
allres = [torch.empty_like(res) for _ in range(world)]
dist.all_gather(allpos, pos)
dist.all_gather(allres, res)
ok = True
if rank == 0:
    P = torch.stack(allpos).cpu().numpy().reshape(world, N, 3)
    R = torch.stack(allres).cpu().numpy().astype(bool)
    assert R.any(axis=0).all(), "some vertex is resident on no rank"
    merged = np.zeros((N, 3), np.float32)
    for r in range(world):
        merged[R[r]] = P[r][R[r]]
    # replicas of shared vertices must be bit-identical
    shared = R.sum(axis=0) > 1
    for r in range(world):
        sel = R[r] & shared
        if not np.array_equal(P[r][sel].view(np.uint32), merged[sel].view(np.uint32)):
            ok = False
            print("FAIL: rank %d replicas differ from the merged copy" % r)
    single = ts.SoftBody(v, t, None, pp, solver="jacobi", iters=a.iters, cluster_size=a.cluster_size, device=local)
    for _ in range(a.substeps // 20):
        single.step(pp)
    ref = single.pos.reshape(N, 3).astype(np.float64)
    err = float(np.max(np.linalg.norm(merged - ref, axis=1) / np.linalg.norm(ref, axis=1)))
    touched = bool(np.any(ref[:, 1] == 0.0))
    print("world %d: tets/rank %d, tiles %d (%d boundary), boundary verts %d (%.1f KB all-reduced per iteration), shared %d, "
          "vector-rel err vs 1 GPU %.3e, floor contact %s" % (world, info["localTets"], info["numClusters"], info["boundaryTiles"], info["boundaryVerts"],
                                                              info["boundaryVerts"] * 16 / 1024, int(shared.sum()), err, touched))
    if not (err <= 1e-5):
        ok = False
        print("FAIL: N-rank result differs from the single-GPU result")
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
body.synchronize()
dist.barrier()   # every rank idle before any exchange buffer is freed
body.close()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
